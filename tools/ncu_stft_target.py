"""ncu target: one forward of the general-kernel STFT (tcgen05 GEMM) at N=256 and N=4096 sequences of 300 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skeleton_action_recognition_b200.layers.virtual_radar import _STFTKernels
k = _STFTKernels(256, 16, True, "cuda:0")
for N in (256, 4096):
    iq = torch.randn(N, 300, 2, device="cuda")
    with torch.no_grad():
        for _ in range(2): k.logmag(iq)
torch.cuda.synchronize()
