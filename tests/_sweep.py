"""Tuning sweep (not a test): python tests/_sweep.py <lib> <warps> <ctas_per_sm> [stages]"""
import os, sys
lib, W, C = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
S = int(sys.argv[4]) if len(sys.argv) > 4 else 0
os.environ["VR_B200_LIB_OVERRIDE"] = os.path.abspath(lib)
sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
_cabi.lib().vr_set_tuning(W, C, S)
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
out = []
for N in (256, 4096):
    nb = max(2, int(300e6 // (N * 180000)) + 1)
    xs = [torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3 for _ in range(nb)]
    for i in range(3): layer(xs[i % nb])
    torch.cuda.synchronize()
    K = 30
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for i in range(K):
        layer(xs[i % nb]); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(K))
    out.append("N=%d %.1fus %.2fM/s" % (N, ts[K // 2] * 1e3, N / ts[K // 2] / 1e3))
    del xs
src, dst = layer.src, layer.dst
pl = _cabi.plan(256, 300, 25, 2, src, dst)
print(os.path.basename(lib), "W=%d C=%d S=%d(smem %d)" % (W, pl["ctas_per_sm"], pl["ring_stages"], pl["smem_bytes"]), " | ".join(out), flush=True)
