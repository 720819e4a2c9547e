"""Timing aid (not a test): forward_host (pinned host in -> host out) for several sub-batch sizes, N=256 NTU."""
import sys, time
sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
xh = [(torch.randn(256, 3, 300, 25, 2) * 0.3).pin_memory() for _ in range(2)]
oh = torch.empty(256, 256, 19).pin_memory()
# raw copy speed for reference
xd = torch.empty_like(xh[0], device='cuda')
for _ in range(3): xd.copy_(xh[0], non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): xd.copy_(xh[i % 2], non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("plain H2D of one batch: %.3f ms = %.1f GB/s" % (dt * 1e3, xd.numel() * 4 / dt / 1e9))
for sb in (0, 256, 128, 86, 64, 43, 32, 16, 8):
    for _ in range(3): layer.forward_host(xh[0], out=oh, sub_batch=sb)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = 50
    for i in range(K): layer.forward_host(xh[i % 2], out=oh, sub_batch=sb)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print("sub_batch=%3d  %.3f ms/step  %.0f seq/s" % (sb, dt * 1e3, 256 / dt), flush=True)
