"""GPU parity at FULL size and in the regimes the crops do not reach (through the C ABI; nothing reads /root/reference):

  * BASELINE configs 1 and 3 -- notebook cells 4 / 2 / 3: NTU x550 (T = 165 000, lambda 9e-4), CMU x20 (T = 55 020,
    41 bones, lambda 5e-3), simulated gait x10 (T = 81 920, 16 bones, lambda 5e-4: ranges up to 6 m, phases up to
    151 000 rad) with the notebook's coordinate-innermost strides.  Checked against (a) the REAL reference's outputs
    committed by tests/golden/make_golden_full.py (every 48th spectrogram column, every 16th baseband sample, all row
    checksums), (b) the oracle run in the same test on the full output, with the tiered criterion of SURVEY 8d,
    (c) BASELINE.md section 3 rows B / C / D (shape, sum, min, max, argmax).
  * far targets: randn + 5..10 m offset at lambda 5e-4 (phases 1.2e5 .. 2.6e5 rad, beyond libdevice's fast range).
  * NaN contract: NaN in => NaN out on exactly the frames the reference marks; |u| > 1 (acos -> NaN in the reference,
    layers/virtual_radar.py:104-105) gives NaN on exactly the same baseband samples.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import virtual_radar_oracle as vro
from tests import fixtures as fx

pytestmark = pytest.mark.gpu

REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _layer(**kw):
    from skeleton_action_recognition_b200 import VirtualRadar
    return VirtualRadar(device="cuda:0", **kw).to("cuda:0")


def _record(name, rep):
    try:
        os.makedirs(REPORT, exist_ok=True)
        with open(os.path.join(REPORT, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **rep}) + "\n")
    except OSError:
        pass


@pytest.mark.parametrize("name", sorted(fx.FULL))
def test_notebook_configs_at_full_size(name):
    x, kw, gold = fx.full_case(name)
    assert tuple(x.stride()) == gold["x_strides"] and x.stride(1) == 1          # the notebook's layout: range mode "fma"
    layer = _layer(**kw)
    xg = x.cuda()
    assert xg.stride() == x.stride()
    out, iq = layer.forward_debug(xg)
    torch.cuda.synchronize()
    out, iq = out.cpu().numpy(), iq.cpu().numpy()
    ka = json.load(open(os.path.join(fx.GOLDEN, "known_answers.json")))[fx.FULL[name]["row"]]
    assert list(out.shape) == ka["shape"]

    # largest phase on the path (the regime this test exists for)
    d = torch.linalg.vector_norm(x.double(), dim=1).max().item()
    theta_max = 4 * np.pi * d / kw["wavelength"]
    # (a) the real reference, sub-sampled
    s = gold["stride"]
    rep_ref = vro.parity_report(out[:, :, ::s], gold["y"])
    rms = np.sqrt((gold["iq"].astype(np.float64) ** 2).sum(-1).mean())
    iq_err = np.abs(iq[:, ::16].astype(np.float64) - gold["iq"]) / rms
    rowsum = out.astype(np.float64).sum(axis=2)
    # (b) the oracle on the whole output
    ref = vro.forward(x, distance=vro.distance_mode_for(x), **kw).numpy()
    rep = vro.parity_report(out, ref)
    rep.update(theta_max_rad=theta_max, iq_max_rel_rms=float(iq_err.max()), iq_median_rel_rms=float(np.median(iq_err)),
               vs_reference_subsample={k: rep_ref[k] for k in ("global_abs_over_peak", "t1", "t2")},
               rowsum_max_abs_diff=float(np.abs(rowsum - gold["rowsum"]).max()),
               sum=float(out.astype(np.float64).sum()), sum_ref=ka["sum"], max=float(out.max()), min=float(out.min()))
    _record("fullsize/" + name, rep)
    assert np.array_equal(ref[:, :, ::s], gold["y"])                  # the oracle IS the reference here, bit for bit
    assert vro.parity_ok(rep_ref), rep_ref
    assert vro.parity_ok(rep), rep
    assert np.median(iq_err) < 2e-6 and iq_err.max() < 1e-3, (np.median(iq_err), iq_err.max())
    # (c) BASELINE.md rows B / C / D
    assert abs(out.max() - ka["max"]) < 1e-4
    assert [int(i) for i in np.unravel_index(np.argmax(out), out.shape)] == ka["argmax"]
    # the sum and the minimum are dominated by bins 60-140 dB below the peak, where two float32 evaluation orders differ
    # by design; they are pinned loosely and reported exactly in gpurun_out/parity_report.jsonl
    assert abs(out.astype(np.float64).sum() - ka["sum"]) <= 2e-4 * abs(ka["sum"])
    assert out.min() >= np.log(np.float32(1e-6)) - 1e-3
    if name == "gait":
        assert theta_max > 1.4e5                                      # beyond anything the crops exercise (44 k rad)


@pytest.mark.parametrize("name", sorted(fx.FULL))
def test_notebook_pad_frames_on_the_device(name):
    """`utils.pad_frames` (reference utils.py:82-89: Gaussian along the JOINT axis, cubic in time, float64) on the GPU
    (C ABI vr_pad_frames_joints) at the notebook's full sizes: float32 positions equal to the real reference's (golden
    samples made with the real utils.py) and to scipy run here, in both output layouts; and the two-launch device
    pipeline `forward_notebook` gives the very spectrogram of the host-prepared notebook tensor."""
    from skeleton_action_recognition_b200 import pad_frames_notebook
    raw, k = fx.full_raw(name)
    x_host, kw, gold = fx.full_case(name)                     # scipy on the host + the notebook's transposes
    dev_raw = torch.from_numpy(raw).cuda()
    up = pad_frames_notebook(dev_raw, k)                      # (1, 3, kT, V, 1), the notebook's strides
    assert tuple(up.shape) == tuple(x_host.shape) and up.stride()[1:4] == x_host.stride()[1:4] and up.stride(1) == 1
    got = up.cpu()
    rows = got[0, :, :, :, 0].permute(1, 2, 0).numpy()        # (kT, V, 3)
    assert np.array_equal(rows[::997], gold["up"])            # the real utils.pad_frames, cast like torch.Tensor(...)
    # every position against scipy on this host.  Equal except at zero crossings of a coordinate: there the cubic's terms
    # cancel, |value| is 1e-8 .. 1e-37 of the coefficients, and the two float64 evaluations (scipy's B-spline recursion,
    # the kernel's Horner form) differ by enough to land on the two sides of a float32 rounding boundary -- 1 of 4.2 M
    # positions for the gait file, 4 of 12.4 M for NTU, none for CMU (tools/pad_nb_diag.py prints them).  One float32 ulp
    # of a 1e-8 m coordinate is 1e-15 m: it cannot move a range.
    a, b = got.numpy(), x_host.numpy()
    bad = a != b
    assert bad.sum() <= 1e-6 * a.size, int(bad.sum())
    if bad.any():
        scale = np.abs(b).max()
        assert np.abs(b[bad]).max() < 1e-6 * scale, (np.abs(b[bad]).max(), scale)
        ulp = np.abs(a[bad].view(np.int32).astype(np.int64) - b[bad].view(np.int32).astype(np.int64))
        assert ulp.max() == 1, ulp
    assert abs(float(got.double().sum()) - gold["up_sum"]) < 1e-9 * max(1.0, abs(gold["up_sum"]))
    planar = pad_frames_notebook(dev_raw, k, planar=True)
    assert planar.is_contiguous() and torch.equal(planar, up.contiguous())
    layer = _layer(**kw)
    want = layer(x_host.cuda())                               # the host-prepared notebook tensor
    dev = layer(up)                                           # same strides -> same range rounding mode
    assert torch.equal(layer.forward_notebook(dev_raw, k), dev)       # planar layout + explicit mode: the same bits
    assert not torch.equal(layer(planar), dev)                # the planar copy alone would select the other mode
    if not bad.any():
        assert torch.equal(dev, want)
    assert vro.parity_ok(vro.parity_report(dev.cpu().numpy(), want.cpu().numpy()))
    s = gold["stride"]
    assert vro.parity_ok(vro.parity_report(dev.cpu().numpy()[:, :, ::s], gold["y"]))
    _record("notebook_pad/" + name, {"positions_differing_from_scipy": int(bad.sum()), "positions": int(a.size),
                                     "spectrogram_bit_equal_to_host_prepared": bool(torch.equal(dev, want))})


def test_notebook_pad_frames_shapes_and_errors():
    from skeleton_action_recognition_b200 import pad_frames_notebook
    from oracle.pad_frames import pad_frames as host_pad
    g = torch.Generator().manual_seed(31)
    for shape, k, sigma, dt in (((40, 5, 3), 7, 3, torch.float64), ((4, 3, 3), 3, 1, torch.float32), ((500, 42, 3), 5, 2, torch.float64),
                                ((64, 25, 2), 9, 3, torch.float32)):
        a = torch.randn(*shape, generator=g, dtype=dt)
        want = torch.Tensor(host_pad(a.numpy(), k, sigma))                       # (kT, V, C) float32
        got = pad_frames_notebook(a.cuda(), k, sigma)
        assert tuple(got.shape) == (1, shape[2], k * shape[0], shape[1], 1)
        assert torch.equal(got[0, :, :, :, 0].permute(1, 2, 0).cpu(), want), (shape, k)
    batch = torch.randn(3, 50, 6, 3, generator=g, dtype=torch.float64)
    got = pad_frames_notebook(batch.cuda(), 4)
    for n in range(3):
        assert torch.equal(got[n:n + 1], pad_frames_notebook(batch[n].cuda(), 4))
    with pytest.raises(ValueError):
        pad_frames_notebook(torch.zeros(3, 5, 3, device="cuda"), 4)              # T < 4
    with pytest.raises(NotImplementedError):
        pad_frames_notebook(torch.zeros(9000, 2, 3, device="cuda", dtype=torch.float64), 2)
    with pytest.raises(RuntimeError):
        pad_frames_notebook(torch.zeros(10, 5, 3), 4)


@pytest.mark.parametrize("offset,lam", [((0., 0., 5.), 5e-4), ((6., -3., 7.), 5e-4), ((2., 9., -4.), 9e-4)])
def test_far_targets_large_phase(offset, lam):
    """randn bodies 5-10 m from the radar: theta = 4 pi d / lambda reaches 1.2e5 .. 2.6e5 rad, where one ulp of theta
    is 8e-3 .. 1.6e-2 rad and the two-term Cody-Waite reduction has to carry ~16 bits of the quotient."""
    g = torch.Generator().manual_seed(int(offset[0] * 10 + 3))
    x = torch.randn(6, 3, 300, 25, 2, generator=g) * 0.3 + torch.tensor(offset).view(1, 3, 1, 1, 1)
    layer = _layer(wavelength=lam)
    out, iq = layer.forward_debug(x.cuda())
    o = vro.OracleVirtualRadar(wavelength=lam)
    iq_ref = o.iq(x, distance="seq").numpy()
    ref = vro.stft_logmag(torch.from_numpy(iq_ref), o.stft, 256).numpy()
    rep = vro.parity_report(out.cpu().numpy(), ref)
    rms = np.sqrt((iq_ref.astype(np.float64) ** 2).sum(-1).mean())
    err = np.abs(iq.cpu().numpy().astype(np.float64) - iq_ref) / rms
    rep.update(theta_max_rad=float(4 * np.pi * torch.linalg.vector_norm(x, dim=1).max() / lam),
               iq_max_rel_rms=float(err.max()), iq_median_rel_rms=float(np.median(err)))
    _record("far/%s" % (offset,), rep)
    assert rep["theta_max_rad"] > 1.1e5
    assert np.median(err) < 2e-6 and err.max() < 1e-4, (np.median(err), err.max())
    assert vro.parity_ok(rep), rep


def test_nan_in_nan_out():
    """A NaN coordinate poisons the baseband sample of its time step and, through the STFT, exactly the frames whose
    window covers it -- the same set of bins as in the reference; everything else is untouched."""
    x = fx.s1_iid(4, seed=8, shape=(3, 800, 25, 2))
    x[1, 0, 333, 7, 0] = float("nan")              # a source joint
    x[2, 2, 40, 24, 1] = float("nan")              # a joint that is only ever a bone's far end
    x[2, 1, 700:703, 3, 0] = float("inf")
    layer = _layer(wavelength=5e-4)
    out, iq = layer.forward_debug(x.cuda())
    out, iq = out.cpu().numpy(), iq.cpu().numpy()
    o = vro.OracleVirtualRadar(wavelength=5e-4)
    iq_ref = o.iq(x, distance="seq").numpy()
    ref = vro.stft_logmag(torch.from_numpy(iq_ref), o.stft, 256).numpy()
    assert np.array_equal(np.isnan(iq), np.isnan(iq_ref))
    assert np.isnan(iq_ref[1, 333]).all() and np.isnan(iq_ref[2, 40]).all() and np.isnan(iq_ref[2, 700:703]).all()
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.isnan(ref).any() and not np.isnan(ref[0]).any() and not np.isnan(ref[3]).any()
    clean = ~np.isnan(ref).any(axis=1)             # (N, F) frames without NaN
    assert clean[1].sum() > 10 and clean[2].sum() > 5
    for n in range(4):
        rep = vro.parity_report(out[n:n + 1][:, :, clean[n]], ref[n:n + 1][:, :, clean[n]])
        assert vro.parity_ok(rep), (n, rep)


def test_aspect_cosine_beyond_one_gives_nan_like_the_reference():
    """Bones that point exactly at the radar from 12-40 m away: rounding makes |u| = |A.B| / (|A||B| + 1e-6) exceed 1 for
    some of them, the reference's acos returns NaN (layers/virtual_radar.py:104-105) and the sample becomes NaN.  The
    kernel computes u with the reference's rounding, so the same samples -- no more, no fewer -- are NaN."""
    g = torch.Generator().manual_seed(3)
    N, T, V, M = 3, 800, 6, 2
    dirs = torch.randn(N, 3, T, 1, M, generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    r = 12 + 30 * torch.rand(N, 1, T, 1, M, generator=g)
    radial = dirs * (r + torch.arange(V).float().view(1, 1, 1, V, 1) * 0.25)
    x = torch.randn(N, 3, T, V, M, generator=g) * 0.3 + 15
    x[0, :, 300:420] = radial[0, :, 300:420]       # a window of radial bones in sequence 0
    x[2, :, :, :, 1] = radial[2, :, :, :, 1]       # one whole body of sequence 2
    edges = [(i, i + 1) for i in range(V - 1)]
    for xx in (x, x.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)):      # both range rounding modes
        mode = vro.distance_mode_for(xx)
        layer = _layer(edges=edges, wavelength=1e-3)
        out, iq = layer.forward_debug(xx.cuda())
        out, iq = out.cpu().numpy(), iq.cpu().numpy()
        o = vro.OracleVirtualRadar(edges=edges, wavelength=1e-3)
        iq_ref = o.iq(xx, distance=mode).numpy()
        ref = vro.stft_logmag(torch.from_numpy(iq_ref), o.stft, 256).numpy()
        nan_ref = np.isnan(iq_ref[..., 0])
        assert nan_ref[0].sum() >= 3 and nan_ref[1].sum() == 0 and nan_ref[2].sum() >= 3, nan_ref.sum(axis=1)
        assert np.array_equal(np.isnan(iq), np.isnan(iq_ref)), (mode, np.isnan(iq[..., 0]).sum(axis=1), nan_ref.sum(axis=1))
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        clean = ~np.isnan(ref).any(axis=1)
        for n in (0, 1):
            assert clean[n].sum() > 5
            rep = vro.parity_report(out[n:n + 1][:, :, clean[n]], ref[n:n + 1][:, :, clean[n]])
            assert vro.parity_ok(rep), (mode, n, rep)
