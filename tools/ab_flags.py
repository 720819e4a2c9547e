"""Timing aid (not a test): vr_forward_f32 with and without VR_FLAG_INPUTS_READY on rotating independent batches."""
import sys
sys.path.insert(0, '.')
import torch, ctypes
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
L = _cabi.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for N, K in ((256, 1000), (128, 1000), (512, 500), (1024, 300), (16384, 30)):
    nb = max(3, int(320e6 // (N * 199456)) + 1)
    xs = [torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3 for _ in range(nb)]
    outs = [torch.empty(N, 256, 19, device='cuda') for _ in range(nb)]
    ref = [layer(x) for x in xs[:3]]
    for flags in (0, 2):
        def step(i):
            rc = L.vr_forward_f32(xs[i % nb].data_ptr(), N, 300, 25, 2, layer._src_c, layer._dst_c, 24, layer.wavelength.data_ptr(),
                                  layer.radar_location.data_ptr(), 256, 16, flags, outs[i % nb].data_ptr(), st)
            assert rc == 0
        for i in range(nb): step(i)
        torch.cuda.synchronize()
        assert all(torch.equal(outs[i], ref[i]) for i in range(3))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K): step(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("N=%6d flags=%d  %9.2f us/step  %.2f M/s  (%.3f of 6541 GB/s)" % (N, flags, ms * 1e3, N / ms / 1e3, N / ms * 1e3 * 199456 / 6541.1e9), flush=True)
    del xs, outs
