"""Short target for ncu captures: a few launches of the fused kernel at N=256 and one at N=16384."""
import sys; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar
big = len(sys.argv) > 1 and sys.argv[1] == 'big'
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
g = torch.Generator().manual_seed(0)
n = 4096 if big else 256
x = (torch.randn(256, 3, 300, 25, 2, generator=g) * 0.3).cuda().repeat(n // 256, 1, 1, 1, 1)
for _ in range(6):
    y = layer(x)
torch.cuda.synchronize()
print(y.sum().item())
