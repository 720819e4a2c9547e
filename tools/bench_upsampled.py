"""Timing aid (not a test): fused forward_upsampled vs pad_frames -> forward_image, k=250, CUDA events."""
import sys, json
sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar, pad_frames
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
res = {}
for N in (32, 148, 296):
    x = torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3
    buf = torch.empty(N, 3, 75000, 25, 2, device='cuda')
    def fused(): return layer.forward_upsampled(x, 250, 3, image_size=256)
    def fused_spec(): return layer.forward_upsampled(x, 250, 3)
    def unfused(): return layer.forward_image(pad_frames(x, 250, 3, out=buf), 256)
    for name, fn in (("fused_image", fused), ("fused_spectrogram", fused_spec), ("unfused_image", unfused)):
        for i in range(2): fn()
        torch.cuda.synchronize()
        K = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        res["N%d_%s" % (N, name)] = dict(ms=ms, seq_per_s=N / ms * 1e3)
        print(N, name, "%.3f ms" % ms, "%.0f seq/s" % (N / ms * 1e3), flush=True)
    del buf
print(json.dumps(res))
