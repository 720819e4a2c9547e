"""Short target for ncu: the fused resize at N=4096 (NTU shape)."""
import sys; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
g = torch.Generator().manual_seed(0)
x = (torch.randn(256, 3, 300, 25, 2, generator=g) * 0.3).cuda().repeat(16, 1, 1, 1, 1)
for _ in range(5):
    y = layer.forward_image(x, 256)
torch.cuda.synchronize()
print(y.sum().item())
