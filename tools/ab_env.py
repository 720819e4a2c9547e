"""A/B timing aid (not a test): python tools/ab_env.py  -- times the current library at several batch sizes; run it
twice with different environment switches (e.g. VR_B200_STATIC_JOBS=1) to compare."""
import os, sys
sys.path.insert(0, '.')
import torch, ctypes
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
L = _cabi.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
base = (torch.randn(256, 3, 300, 25, 2) * 0.3).cuda()
for N, K in ((256, 500), (1024, 200), (2048, 100), (4096, 60), (16384, 30), (65536, 8)):
    nb = max(2, int(320e6 // (N * 199456)) + 1) if N < 65536 else 1
    xs = [base.repeat(N // 256, 1, 1, 1, 1) for _ in range(nb)]
    outs = [torch.empty(N, 256, 19, device='cuda') for _ in range(nb)]
    def step(i):
        rc = L.vr_forward_f32(xs[i % nb].data_ptr(), N, 300, 25, 2, layer._src_c, layer._dst_c, 24, layer.wavelength.data_ptr(),
                              layer.radar_location.data_ptr(), 256, 16, 0, outs[i % nb].data_ptr(), st)
        assert rc == 0
    for i in range(3): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%s N=%6d  %9.2f us  %.2f M/s  (%.3f of 6541 GB/s)" % (os.environ.get("VR_B200_STATIC_JOBS", "dynamic"), N, ms * 1e3, N / ms / 1e3, N / ms * 1e3 * 199456 / 6541.1e9), flush=True)
    del xs, outs
