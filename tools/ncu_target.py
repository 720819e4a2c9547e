"""Short target for ncu captures (not a test): a few launches of the fused forward.
    python tools/ncu_target.py [n] [sched] [flag]     n sequences (default 256), schedule -1/0/1, flag 1 = VR_FLAG_INPUTS_READY"""
import sys; sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
n = 4096 if (len(sys.argv) > 1 and sys.argv[1] == 'big') else int(sys.argv[1]) if len(sys.argv) > 1 else 256
if len(sys.argv) > 2: _cabi.set_schedule(int(sys.argv[2]))
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
layer.assume_inputs_ready = len(sys.argv) > 3 and sys.argv[3] == '1'
g = torch.Generator().manual_seed(0)
x = (torch.randn(256, 3, 300, 25, 2, generator=g) * 0.3).cuda().repeat((n + 255) // 256, 1, 1, 1, 1)[:n].contiguous()
for _ in range(6):
    y = layer(x)
torch.cuda.synchronize()
print(y.sum().item())
