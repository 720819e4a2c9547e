import sys; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import pad_frames
x = (torch.randn(148, 3, 300, 25, 2) * 0.3).cuda(); o = torch.empty(148, 3, 75000, 25, 2, device="cuda")
for _ in range(3): pad_frames(x, 250, out=o)
torch.cuda.synchronize()
