"""Profiling aid (not a test): per-CTA timeline of one launch.  python tools/timeline.py [N]"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
x = torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3
pl = _cabi.plan(N, 300, 25, 2, layer.src, layer.dst)
buf = torch.zeros(pl['grid'] * 8, dtype=torch.int64, device='cuda')
for _ in range(3): layer(x)
torch.cuda.synchronize()
_cabi.lib().vr_set_timeline_buffer(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); layer(x); e1.record()
torch.cuda.synchronize()
_cabi.lib().vr_set_timeline_buffer(None)
t = buf.cpu().numpy().reshape(-1, 8).astype(np.int64)
t0 = t[:, 0].min()
names = ['entry', 'prologue', 'first_chunk', 'synth_done', 'fft_done', 'exit']
print('event ms %.2f us, grid %d' % (e0.elapsed_time(e1) * 1e3, pl['grid']))
for i, n in enumerate(names):
    v = (t[:, i] - t0) / 1e3
    v = v[t[:, i] > 0]
    if len(v): print("%-12s min %.2f  med %.2f  max %.2f us (%d CTAs)" % (n, v.min(), np.median(v), v.max(), len(v)))
sm = t[:, 7]
cnt = np.bincount(sm, minlength=148)
for k in (1, 2):
    sel = np.isin(sm, np.where(cnt == k)[0])
    if sel.any():
        d = (t[sel, 5] - t[sel, 0]) / 1e3
        print('CTAs on SMs with %d CTA(s): %d, lifetime med %.2f max %.2f us; synth med %.2f fft med %.2f' % (
            k, sel.sum(), np.median(d), d.max(), np.median((t[sel, 3] - t[sel, 2]) / 1e3), np.median((t[sel, 4] - t[sel, 3]) / 1e3)))
