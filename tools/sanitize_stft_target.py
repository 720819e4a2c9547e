"""compute-sanitizer target for the tcgen05 STFT alone (forward + backward, two small shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skeleton_action_recognition_b200.layers.virtual_radar import _STFTKernels
for n_fft, hop, T, N in ((256, 16, 300, 40), (64, 6, 203, 3)):
    k = _STFTKernels(n_fft, hop, True, "cuda:0")
    x = torch.randn(N, T, 2, device="cuda", requires_grad=True)
    k.logmag(x).square().mean().backward()
    torch.cuda.synchronize()
    print("ok", n_fft, float(x.grad.abs().sum()), float(k.wsin.grad.abs().sum()))
