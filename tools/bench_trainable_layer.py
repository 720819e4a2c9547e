"""Timing aid (not a test): the whole layer with trainable STFT kernels (fused baseband launch + tcgen05 STFT) next to the
default layer (one fused launch), forward under no_grad and a training step's forward + backward.
python tools/bench_trainable_layer.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skeleton_action_recognition_b200 import VirtualRadar

def timed(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

g = torch.Generator().manual_seed(0)
base = (torch.randn(256, 3, 300, 25, 2, generator=g) * 0.3).cuda()
for N in (256, 4096):
    x = base.repeat(N // 256, 1, 1, 1, 1).contiguous()
    plain = VirtualRadar(wavelength=5e-4, device="cuda:0").to("cuda:0")
    trained = VirtualRadar(wavelength=5e-4, train_stft_kernel=True, device="cuda:0").to("cuda:0")
    with torch.no_grad():
        t_plain = timed(lambda: plain(x)); t_tr = timed(lambda: trained(x))
    def step():
        trained.zero_grad()
        trained(x).square().mean().backward()
    t_step = timed(step, 10)
    print("N %5d: default layer %.3f ms | trainable STFT kernels: forward %.3f ms, forward + backward (kernel gradients) %.3f ms"
          % (N, t_plain, t_tr, t_step), flush=True)
