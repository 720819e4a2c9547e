import sys; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
x = torch.randn(32, 3, 300, 25, 2, device='cuda') * 0.3
for _ in range(3):
    y = layer.forward_upsampled(x, 250, 3, image_size=256)
torch.cuda.synchronize()
print(y.sum().item())
