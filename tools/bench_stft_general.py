"""Timing aid (not a test): the general-kernel STFT on the tensor cores (tcgen05 GEMM) against the torch / cuBLAS path it
replaced, forward and forward+backward.  python tools/bench_stft_general.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skeleton_action_recognition_b200.layers.virtual_radar import _STFTKernels

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n_fft, hop, T, N in ((256, 16, 300, 256), (256, 16, 300, 4096), (256, 16, 75000, 4), (512, 32, 300, 256), (128, 16, 300, 256)):
    k = _STFTKernels(n_fft, hop, True, "cuda:0")
    iq = torch.randn(N, T, 2, device="cuda")
    F = T // hop + 1
    flops = 2.0 * N * F * (2 * n_fft) * (2 * n_fft)
    with torch.no_grad():
        t_new = timed(lambda: k.logmag(iq)); t_old = timed(lambda: k._logmag_torch(iq))
    x = iq.clone().requires_grad_(True)
    def fb(fn):
        k.zero_grad(); x.grad = None
        fn(x).sum().backward()
    t_new_fb = timed(lambda: fb(k.logmag), 10); t_old_fb = timed(lambda: fb(k._logmag_torch), 10)
    print("n_fft %4d hop %3d T %6d N %5d (%d frames): forward tcgen05 %.3f ms (%.1f TFLOP/s of useful float32 GEMM) | torch+cuBLAS %.3f ms | fwd+bwd tcgen05 %.3f ms | torch %.3f ms"
          % (n_fft, hop, T, N, N * F, t_new, flops / t_new / 1e9, t_old, t_new_fb, t_old_fb), flush=True)
