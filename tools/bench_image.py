"""Timing aid (not a test): fused forward_image vs forward + torch F.interpolate, CUDA events."""
import sys, json
sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
res = {}
for N, T in ((256, 300), (4096, 300), (16, 75000)):
    nb = max(2, int(300e6 // (N * 3 * T * 50 * 4)) + 1)
    xs = [torch.randn(N, 3, T, 25, 2, device='cuda') * 0.3 for _ in range(nb)]
    def fused(i): return layer.forward_image(xs[i % nb], 256)
    def unfused(i): return torch.nn.functional.interpolate(layer(xs[i % nb]).unsqueeze(1), 256)
    def plain(i): return layer(xs[i % nb])
    for name, fn in (("fused", fused), ("unfused", unfused), ("spectrogram_only", plain)):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        K = 50 if N * T < 1e6 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K): fn(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        res["N%d_T%d_%s" % (N, T, name)] = dict(ms=ms, seq_per_s=N / ms * 1e3)
        print(N, T, name, "%.3f ms" % ms, "%.0f seq/s" % (N / ms * 1e3), flush=True)
    del xs
print(json.dumps(res))
