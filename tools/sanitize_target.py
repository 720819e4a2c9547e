"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel variant family once."""
import sys
sys.path.insert(0, '.')
import torch
from skeleton_action_recognition_b200 import VirtualRadar, pad_frames
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
g = torch.Generator().manual_seed(0)
x = (torch.randn(3, 3, 300, 25, 2, generator=g) * 0.3).cuda()
y = layer(x)                                            # short, TMA ring, bulk store
img = layer.forward_image(x, 64)                        # fused resize
xl = (torch.randn(1, 3, 5000, 25, 2, generator=g) * 0.3).cuda()
yl = layer(xl)                                          # long: several jobs, parked sums
il = layer.forward_image(xl, 96)                        # sparse frames
xr = (torch.randn(2, 3, 40, 25, 2, generator=g) * 0.3).cuda()
yu = layer.forward_upsampled(xr, 20, 3, image_size=64)  # spline + team-evaluated chunks
up = pad_frames(xr, 20, 3)
xo = (torch.randn(2, 3, 301, 17, 1, generator=g) * 0.3).cuda()      # unaligned, odd M, generic V*M
yo = VirtualRadar(edges=[(i, i + 1) for i in range(16)], wavelength=1e-3, radar_location=[0.1, 0.2, 0.3], device='cuda:0').to('cuda:0')(xo)
big = x.repeat(120, 1, 1, 1, 1)                         # 360 jobs > 296 CTAs: dynamic scheduling
yb = layer(big)
t = VirtualRadar(wavelength=5e-3, train_wavelength=True, train_radar_location=True, device='cuda:0').to('cuda:0')
xg = x.clone().requires_grad_(True)
t(xg).square().mean().backward()                        # adjoint kernels
# round 2: team-job schedule, notebook up-sampling, tcgen05 STFT (forward + backward)
from skeleton_action_recognition_b200 import _cabi, pad_frames_notebook
_cabi.set_schedule(1)
yt = layer(big)                                         # team-job kernel: 360 jobs on 296 x 2 teams + tickets
yt2 = layer(x)                                          # fewer jobs than teams
_cabi.set_schedule(-1)
assert torch.equal(yt, yb) and torch.equal(yt2, y)
nb = pad_frames_notebook((torch.randn(50, 17, 3, generator=g, dtype=torch.float64)).cuda(), 6)
yn = layer17 = None
k = VirtualRadar(wavelength=5e-3, train_stft_kernel=True, device='cuda:0').to('cuda:0')
xk = x.clone().requires_grad_(True)
k(xk).square().mean().backward()                        # tcgen05 GEMM forward + both backward GEMMs + synthesis adjoint
torch.cuda.synchronize()
print("ok", float(y.sum()), float(img.sum()), float(yl.sum()), float(il.sum()), float(yu.sum()), float(yo.sum()), float(yb.sum()), float(t.wavelength.grad))
