"""GPU tests of the fused consumer stage: VirtualRadar.forward_image through the C ABI
(vr_forward_image_f32) must equal `F.interpolate(layer(x).unsqueeze(1), image_size)` -- reference
models/resnet.py:24-26 -- bit for bit, against both the CPU restatement (oracle/resize.py) applied to the
unfused launch and torch's own CUDA interpolate.  The unfused launch itself is pinned to the reference by
tests/test_parity_gpu.py.  Nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

from oracle import resize
from oracle import virtual_radar_oracle as vro
from tests import fixtures as fx

pytestmark = pytest.mark.gpu


def _layer(**kw):
    from skeleton_action_recognition_b200 import VirtualRadar
    return VirtualRadar(device="cuda:0", **kw).to("cuda:0")


def _check(layer, x, size):
    xg = x.cuda()
    spec = layer(xg)
    img = layer.forward_image(xg, size)
    torch.cuda.synchronize()
    assert tuple(img.shape) == (x.shape[0], 1, size, size)
    want = resize.resize_nearest(spec.cpu().numpy(), size)
    got = img.cpu().numpy()
    assert np.array_equal(got, want), (size, float(np.abs(got - want).max()), np.argwhere(got != want)[:4])
    assert torch.equal(img, torch.nn.functional.interpolate(spec.unsqueeze(1), size))
    return img


@pytest.mark.parametrize("size", [256, 224, 64, 300, 37, 512, 19, 1])
def test_ntu_shape_all_frames_replicated(size):
    """T=300 -> 19 frames; every frame is kept and replicated (or decimated for size < 19)."""
    _check(_layer(wavelength=5e-4), fx.s1_iid(5), size)


def test_against_the_oracle_end_to_end():
    """oracle forward -> oracle resize vs the fused launch, tiered parity criterion on the image."""
    x = fx.s1_iid(6, seed=8)
    img = _layer(wavelength=5e-4).forward_image(x.cuda(), 256).cpu().numpy()
    ref = resize.resize_nearest(vro.forward(x, wavelength=5e-4, distance="seq").numpy(), 256)
    rep = vro.parity_report(img[:, 0], ref[:, 0])
    assert vro.parity_ok(rep), rep


@pytest.mark.parametrize("T,size", [(3000, 256), (4095, 256), (4096, 256), (6400, 256), (20000, 256), (75000, 256),
                                    (20000, 100), (9000, 300), (3000, 64), (5000, 510)])
def test_long_sequences_several_jobs(T, size):
    """Dense (frames <= columns) and sparse (only the kept frames are transformed) regimes, several jobs
    per sequence, job boundaries inside replicated frames."""
    g = torch.Generator().manual_seed(T + size)
    x = torch.randn(2, 3, T, 25, 2, generator=g) * 0.3
    _check(_layer(wavelength=1e-3), x, size)


@pytest.mark.parametrize("shape,hop,size", [((3, 3, 700, 17, 1), 8, 256), ((2, 3, 129, 5, 3), 16, 256),
                                            ((2, 3, 1001, 42, 1), 100, 128), ((1, 3, 2500, 25, 4), 32, 77)])
def test_other_shapes_and_hops(shape, hop, size):
    g = torch.Generator().manual_seed(shape[2])
    x = torch.randn(*shape, generator=g) * 0.4
    V = shape[3]
    edges = [(i, i + 1) for i in range(V - 1)]
    _check(_layer(edges=edges, wavelength=2e-3, hop_length=hop, radar_location=[0.1, 0.2, -0.3]), x, size)


def test_full_batch_and_persistent_loop():
    x = fx.s1_iid(64)
    layer = _layer(wavelength=5e-4)
    big = x.repeat(12, 1, 1, 1, 1).cuda()                    # 768 sequences: more jobs than resident CTAs
    img = layer.forward_image(big, 256)
    one = _check(layer, x, 256)
    assert torch.equal(img, one.repeat(12, 1, 1, 1))


def test_image_errors():
    layer = _layer(wavelength=5e-4)
    x = torch.zeros(1, 3, 300, 25, 2, device="cuda")
    with pytest.raises(ValueError):
        layer.forward_image(x, 0)
    with pytest.raises(ValueError):
        layer.forward_image(x, 5000)
    with pytest.raises(RuntimeError):
        layer.forward_image(x.cpu(), 256)


def test_empty_batch_and_strided_input():
    layer = _layer(wavelength=5e-4)
    e = torch.zeros(0, 3, 300, 25, 2, device="cuda")
    assert tuple(layer(e).shape) == (0, 256, 19)
    assert tuple(layer.forward_image(e, 64).shape) == (0, 1, 64, 64)
    assert tuple(layer.forward_upsampled(torch.zeros(0, 3, 40, 25, 2, device="cuda"), 10, 3).shape) == (0, 256, 26)
    # coordinate axis innermost (notebook-style strides): the image path must pick the same rounding mode as forward
    g = torch.Generator().manual_seed(77)
    x = (torch.randn(2, 400, 25, 1, 3, generator=g) * 0.5).permute(0, 4, 1, 2, 3).cuda()
    assert x.stride(1) == 1
    assert torch.equal(layer.forward_image(x, 96), torch.nn.functional.interpolate(layer(x).unsqueeze(1), 96))
