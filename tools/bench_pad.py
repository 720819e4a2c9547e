"""Profiling aid (not a test): throughput of the GPU pad_frames pre-stage.  python tools/bench_pad.py [N] [k]"""
import sys, time; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import pad_frames
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
k = int(sys.argv[2]) if len(sys.argv) > 2 else 250
x = torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3
out = torch.empty(N, 3, 300 * k, 25, 2, device='cuda')
for _ in range(2): pad_frames(x, k, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = 5
e0.record()
for _ in range(R): pad_frames(x, k, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
gb = out.numel() * 4 / 1e9
print("N=%d k=%d: %.3f ms, %.1f samples/s, %.1f GB/s written (%.2f of 6531)" % (N, k, ms, N / ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / 6531))
