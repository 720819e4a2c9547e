"""Temporal up-sampling pre-stage.  CPU: the oracle restatement equals the real reference's
`Dataset.pad_frames` + cast on the committed golden vectors (tests/golden/make_golden_pad_frames.py).  GPU: the
CUDA kernel (through the C ABI) equals the golden vectors / the oracle: float32 positions BIT-EQUAL (a float32 rounding
tie of the float64 result could in principle differ between scipy's B-spline evaluation and the kernel's Horner form; none
occurs on any input here, so the tests demand equality and print the ulp statistics if it ever fails), and the up-sampled
batch pushed through VirtualRadar meets the spectrogram parity criterion."""
import os

import numpy as np
import pytest
import torch

from oracle import pad_frames as opf
from tests import fixtures as fx


def _golden():
    z = np.load(os.path.join(fx.GOLDEN, "pad_frames_dataset.npz"))
    g = torch.Generator().manual_seed(int(z["seed_rand"]))
    x_rand = (torch.randn(3, 3, 64, 5, 3, generator=g) * 0.4).numpy()
    return {"ntu": (z["x_ntu"], int(z["k_ntu"]), z["y_ntu"], 1),
            "rand": (x_rand, int(z["k_rand"]), z["y_rand"], int(z["stride_rand"])),
            "short": (z["x_short"], int(z["k_short"]), z["y_short"], 1)}


@pytest.mark.parametrize("name", ["ntu", "rand", "short"])
def test_oracle_equals_reference_dataset_pad_frames(name):
    x, k, y, stride = _golden()[name]
    got = np.stack([opf.dataset_getitem(s, k).numpy() for s in x])[:, :, ::stride]
    assert got.dtype == np.float32 and got.shape == y.shape
    assert np.array_equal(got, y)


def _ulp_report(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7fffffff), ai)
    bi = np.where(bi < 0, -(bi & 0x7fffffff), bi)
    d = np.abs(ai - bi)
    return float((d == 0).mean()), int(d.max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ntu", "rand", "short"])
def test_gpu_pad_frames_equals_reference(name):
    from skeleton_action_recognition_b200 import pad_frames
    x, k, y, stride = _golden()[name]
    out = pad_frames(torch.from_numpy(x).cuda(), num_pad_frames=k).cpu().numpy()[:, :, ::stride]
    assert out.shape == y.shape
    frac, worst = _ulp_report(out, y)
    assert frac == 1.0 and worst == 0, (frac, worst)
    # single-sample form, like Dataset.pad_frames(data)
    one = pad_frames(torch.from_numpy(x[0]).cuda(), num_pad_frames=k).cpu().numpy()[:, ::stride]
    assert np.array_equal(one, out[0])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,k,sigma", [((2, 3, 300, 25, 2), 25, 3), ((1, 3, 1000, 17, 1), 7, 3), ((2, 3, 40, 42, 1), 11, 2),
                                           ((1, 3, 4, 3, 2), 5, 1), ((1, 3, 3000, 4, 1), 3, 3)])
def test_gpu_pad_frames_equals_oracle(shape, k, sigma):
    """More shapes: column blocks (V*M = 42, 50 at T=300 / 1000), the T=4 minimum, another sigma."""
    from skeleton_action_recognition_b200 import pad_frames
    g = torch.Generator().manual_seed(shape[2] + k)
    x = torch.randn(*shape, generator=g) * 0.5 + torch.linspace(0, 2, shape[2])[None, None, :, None, None]
    ref = torch.stack([opf.dataset_getitem(s.numpy(), k, sigma) for s in x]).numpy()
    out = pad_frames(x.cuda(), num_pad_frames=k, sigma=sigma).cpu().numpy()
    frac, worst = _ulp_report(out, ref)
    assert frac == 1.0 and worst == 0, (frac, worst)


@pytest.mark.gpu
def test_gpu_pad_frames_then_virtual_radar_end_to_end():
    """The training input path of the reference (utils.py:128-140 -> models/resnet.py:24): up-sample, then the
    layer.  GPU pre-stage + GPU layer against oracle pre-stage + oracle layer."""
    from oracle import virtual_radar_oracle as vro
    from skeleton_action_recognition_b200 import VirtualRadar, pad_frames
    x = fx.s3_smooth(2, T=60)
    k = 40
    up_ref = torch.stack([opf.dataset_getitem(s.numpy(), k) for s in x])
    ref = vro.forward(up_ref, wavelength=5e-3, distance="seq").numpy()
    layer = VirtualRadar(wavelength=5e-3, device="cuda:0").to("cuda:0")
    out = layer(pad_frames(x.cuda(), num_pad_frames=k)).cpu().numpy()
    assert out.shape == ref.shape == (2, 256, 60 * k // 16 + 1)
    assert vro.parity_ok(vro.parity_report(out, ref))


@pytest.mark.gpu
def test_gpu_pad_frames_errors():
    from skeleton_action_recognition_b200 import pad_frames
    with pytest.raises(ValueError, match="at least 4 frames"):
        pad_frames(torch.zeros(1, 3, 3, 5, 1, device="cuda"))
    with pytest.raises(ValueError):
        pad_frames(torch.zeros(1, 3, 30, 5, 1, device="cuda", dtype=torch.float64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        pad_frames(torch.zeros(1, 3, 30, 5, 1))
    with pytest.raises(NotImplementedError, match="too long"):
        pad_frames(torch.zeros(1, 3, 20000, 5, 1, device="cuda"), num_pad_frames=2)
