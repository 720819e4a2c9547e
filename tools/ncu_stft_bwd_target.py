"""ncu target: one forward + backward of the general-kernel STFT (tcgen05 GEMMs) at N=4096 sequences of 300 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skeleton_action_recognition_b200.layers.virtual_radar import _STFTKernels
k = _STFTKernels(256, 16, True, "cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.randn(N, 300, 2, device="cuda", requires_grad=True)
for _ in range(2):
    k.zero_grad(); x.grad = None
    k.logmag(x).sum().backward()
torch.cuda.synchronize()
