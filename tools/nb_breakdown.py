import sys; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar, pad_frames_notebook, _cabi
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for name, T, V, k, dt, E in (("ntu", 300, 25, 550, torch.float32, None), ("cmu", 2751, 42, 20, torch.float64, [(i, i + 1) for i in range(41)]), ("gait", 8192, 17, 10, torch.float64, [(i, i+1) for i in range(16)])):
    raw = (torch.randn(T, V, 3, dtype=torch.float64) * 0.3).to(dt).cuda()
    kw = dict(wavelength=9e-4, device="cuda:0")
    if E: kw["edges"] = E
    lay = VirtualRadar(**kw).to("cuda:0")
    x = pad_frames_notebook(raw, k, planar=True)
    t_pad = timed(lambda: pad_frames_notebook(raw, k, planar=True))
    t_rad = timed(lambda: lay._run(x, x, 1))
    src, dst = lay.src, lay.dst
    pl = _cabi.plan(1, T * k, V, 1, src, dst)
    print(name, "pad %.3f ms  radar %.3f ms  plan grid %d jobs/seq %d frames/job %d S %d smem %d" % (t_pad, t_rad, pl["grid"], pl["jobs_per_seq"], pl["frames_per_job"], pl["ring_stages"], pl["smem_bytes"]))
