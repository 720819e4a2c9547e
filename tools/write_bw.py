"""Microbenchmark (not a test): write-only and copy bandwidth of this GPU with torch kernels, CUDA events."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device='cuda')      # 4 GiB
y = torch.empty_like(x)
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: x.zero_())
print("write-only (zero_):  %.3f ms  %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
ms = timed(lambda: x.fill_(1.5))
print("write-only (fill_):  %.3f ms  %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
ms = timed(lambda: y.copy_(x))
print("copy (read+write):   %.3f ms  %.0f GB/s" % (ms, 2 * x.numel() * 4 / ms / 1e6))
ms = timed(lambda: x.sum())
print("read-only (sum):     %.3f ms  %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
