"""Tuning sweep (not a test): python tools/sweep.py <lib> [warps ctas_per_sm stages]  -- device-timed via bench-like loop"""
import os, sys, ctypes
lib = sys.argv[1]
W, C, S = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (0, 0, 0)
os.environ["VR_B200_LIB_OVERRIDE"] = os.path.abspath(lib)
sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
L = _cabi.lib()
L.vr_set_tuning(W, C, S)
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
out = []
for N in (256, 16384):
    nb = max(2, int(320e6 // (N * 199456)) + 1)
    xs = [torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3 for _ in range(nb)]
    os_ = [torch.empty(N, 256, 19, device='cuda') for _ in range(nb)]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def step(i):
        rc = L.vr_forward_f32(xs[i % nb].data_ptr(), N, 300, 25, 2, layer._src_c, layer._dst_c, 24, layer.wavelength.data_ptr(),
                              layer.radar_location.data_ptr(), 256, 16, 0, os_[i % nb].data_ptr(), st)
        assert rc == 0
    K = 400 if N == 256 else 30
    for i in range(5): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out.append("N=%d %.2fus %.2fM/s" % (N, ms * 1e3, N / ms / 1e3))
    del xs, os_
pl = _cabi.plan(256, 300, 25, 2, layer.src, layer.dst)
print(os.path.basename(lib), "W=%d C=%d S=%d smem=%d" % (pl["block"] // 32, pl["ctas_per_sm"], pl["ring_stages"], pl["smem_bytes"]), " | ".join(out), flush=True)
