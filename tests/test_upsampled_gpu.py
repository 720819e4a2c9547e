"""GPU tests of the fused temporal up-sampling (north-star stage 1 in front of stages 2-3):
VirtualRadar.forward_upsampled through the C ABI (vr_forward_upsampled_f32) must equal the two-launch path
pad_frames -> forward / forward_image bit for bit (each of those is pinned to the reference separately:
tests/test_pad_frames.py against scipy golden vectors, tests/test_parity_gpu.py against the real layer),
and the CPU oracle chain scipy pad_frames -> oracle forward within the layer's parity criterion."""
import numpy as np
import pytest
import torch

from oracle import pad_frames as opf
from oracle import virtual_radar_oracle as vro
from tests import fixtures as fx

pytestmark = pytest.mark.gpu


def _layer(**kw):
    from skeleton_action_recognition_b200 import VirtualRadar
    return VirtualRadar(device="cuda:0", **kw).to("cuda:0")


@pytest.mark.parametrize("T,k,shape_vm", [(300, 4, (25, 2)), (60, 250, (25, 2)), (64, 33, (25, 2)), (150, 7, (25, 2)),
                                          (40, 20, (17, 1)), (50, 16, (5, 3)), (33, 1, (25, 2)), (200, 3, (42, 1)),
                                          (80, 50, (9, 4))])
def test_equals_two_launch_path(T, k, shape_vm):
    from skeleton_action_recognition_b200 import pad_frames
    V, M = shape_vm
    if k * T <= 128:
        pytest.skip("too short for the STFT")
    g = torch.Generator().manual_seed(T * 1000 + k)
    x = (torch.randn(3, 3, T, V, M, generator=g) * 0.3).cuda()
    edges = [(i, i + 1) for i in range(V - 1)] if V != 25 else None
    layer = _layer(wavelength=1e-3, **({"edges": edges} if edges else {}))
    up = pad_frames(x, k, 3)
    want = layer(up)
    got = layer.forward_upsampled(x, k, 3)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    assert torch.equal(got, want), float((got - want).abs().max())
    for size in (256, 100):
        assert torch.equal(layer.forward_upsampled(x, k, 3, image_size=size), layer.forward_image(up, size))


def test_off_axis_radar_and_sigma():
    from skeleton_action_recognition_b200 import pad_frames
    x = fx.s3_smooth(2, T=120).cuda()
    layer = _layer(wavelength=2e-3, radar_location=[0.3, -0.2, 1.5], hop_length=32)
    for sigma in (1.0, 3, 5.5):
        assert torch.equal(layer.forward_upsampled(x, 25, sigma), layer(pad_frames(x, 25, sigma)))


def test_against_the_cpu_oracle_chain():
    """scipy gaussian_filter1d + interp1d (utils.py:128-140 restated) -> oracle VirtualRadar."""
    x = fx.s3_smooth(3, T=90)
    k = 30
    up = torch.stack([opf.dataset_getitem(x[i].numpy(), k, 3) for i in range(x.shape[0])])
    ref = vro.forward(up, wavelength=5e-4, distance="seq").numpy()
    got = _layer(wavelength=5e-4).forward_upsampled(x.cuda(), k, 3).cpu().numpy()
    rep = vro.parity_report(got, ref)
    assert vro.parity_ok(rep), rep


def test_many_sequences_persistent_loop():
    from skeleton_action_recognition_b200 import pad_frames
    x = fx.s1_iid(40, shape=(3, 40, 25, 2)).cuda()
    layer = _layer(wavelength=5e-4)
    big = x.repeat(10, 1, 1, 1, 1)                                # 400 sequences x several jobs each
    assert torch.equal(layer.forward_upsampled(big, 100, 3, image_size=64), layer.forward_image(pad_frames(big, 100, 3), 64))


def test_upsampled_errors():
    layer = _layer(wavelength=5e-4)
    x = torch.zeros(1, 3, 3, 25, 2, device="cuda")
    with pytest.raises(ValueError):
        layer.forward_upsampled(x, 250, 3)                        # T < 4
    with pytest.raises(ValueError):
        layer.forward_upsampled(torch.zeros(1, 3, 30, 25, 2, device="cuda"), 0, 3)
    with pytest.raises(ValueError):
        layer.forward_upsampled(torch.zeros(1, 3, 30, 25, 2, device="cuda"), 2, 3)   # 60 frames <= n_fft/2
