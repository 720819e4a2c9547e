"""GPU tests of the team-job schedule (vr_team_kernel, C ABI vr_set_schedule): it must produce the SAME BITS as the
cooperative schedule (vr_fused_kernel) -- the arithmetic and the order of every sum are shared -- for every kernel
variant, batch size (fewer jobs than teams, tails, many jobs per team), with overlapping launches and on several streams."""
import numpy as np
import pytest
import torch

from oracle import virtual_radar_oracle as vro
from tests import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.fixture()
def sched():
    from skeleton_action_recognition_b200 import _cabi
    yield _cabi.set_schedule
    _cabi.set_schedule(-1)


def _layer(**kw):
    from skeleton_action_recognition_b200 import VirtualRadar
    return VirtualRadar(device="cuda:0", **kw).to("cuda:0")


def _both(layer, x, sched):
    sched(0)
    a = layer(x)
    sched(1)
    b = layer(x)
    torch.cuda.synchronize()
    return a, b


@pytest.mark.parametrize("n", [1, 2, 3, 7, 255, 256, 592, 593, 1185, 2500])
def test_team_schedule_bit_identical_ntu(n, sched):
    x = fx.s1_iid(64, seed=5).repeat((n + 63) // 64, 1, 1, 1, 1)[:n].cuda()
    x[n // 2:] = x[n // 2:] * 1.25                      # not a pure repetition
    layer = _layer(wavelength=5e-4)
    a, b = _both(layer, x, sched)
    assert torch.equal(a, b)
    if n <= 7:
        ref = vro.forward(x.cpu(), wavelength=5e-4, distance="seq").numpy()
        assert vro.parity_ok(vro.parity_report(b.cpu().numpy(), ref))


@pytest.mark.parametrize("shape,E,loc", [((9, 3, 300, 25, 1), 24, (0., 0., 0.)),          # odd M: scalar twin, generic V*M
                                        ((5, 3, 160, 17, 2), 16, (0.1, 0.2, -0.3)),      # generic V*M, radar off the origin
                                        ((6, 3, 304, 8, 4), 7, (0., 0., 0.)),            # M = 4, 20 frames: the tile is larger than a chunk
                                        ((4, 3, 129, 12, 2), 5, (0., 0., 1.0))])         # minimum T: 5 chunks, last one 1 step
def test_team_schedule_bit_identical_shapes(shape, E, loc, sched):
    g = torch.Generator().manual_seed(shape[2])
    x = (torch.randn(*shape, generator=g) * 0.4).cuda()
    V = shape[3]
    edges = [(i % V, (i * 3 + 1) % V) for i in range(E)]
    edges = [(a, b if b != a else (a + 1) % V) for a, b in edges]
    layer = _layer(edges=edges, wavelength=2e-3, radar_location=list(loc))
    a, b = _both(layer, x, sched)
    assert torch.equal(a, b)
    ref = vro.forward(x.cpu(), edges=edges, wavelength=2e-3, radar_location=loc, distance="seq").numpy()
    assert vro.parity_ok(vro.parity_report(b.cpu().numpy(), ref))


def test_team_schedule_fma_range_mode(sched):
    g = torch.Generator().manual_seed(12)
    x = (torch.randn(6, 300, 25, 2, 3, generator=g) * 0.5).permute(0, 4, 1, 2, 3).cuda()     # coordinate axis innermost
    assert x.stride(1) == 1
    layer = _layer(wavelength=5e-4)
    a, b = _both(layer, x, sched)
    assert torch.equal(a, b)
    assert not torch.equal(b, layer(x.contiguous()))


def test_team_schedule_overlapping_launches_and_streams(sched):
    """Back-to-back launches with VR_FLAG_INPUTS_READY (reads overlap the previous kernel) and launches on two streams:
    the ticket counters, the per-team rings and the tiles must not interfere."""
    layer = _layer(wavelength=5e-4)
    batches = [fx.s1_iid(n, seed=50 + i).cuda() for i, n in enumerate((1300, 256, 5, 2000, 700))]
    sched(0)
    want = [layer(b) for b in batches]
    torch.cuda.synchronize()
    sched(1)
    layer.assume_inputs_ready = True
    for _ in range(4):
        got = [layer(b) for b in batches]
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(got, want))
    layer.assume_inputs_ready = False
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for _ in range(6):
        with torch.cuda.stream(sa):
            ya = layer(batches[0])
        with torch.cuda.stream(sb):
            yb = layer(batches[3])
        outs.append((ya, yb))
    torch.cuda.synchronize()
    for ya, yb in outs:
        assert torch.equal(ya, want[0]) and torch.equal(yb, want[3])


def test_automatic_schedule_switches_by_batch_size(sched):
    """Default (-1): small batches run the cooperative kernel, large ones the team-job kernel; same bits either way."""
    layer = _layer(wavelength=5e-4)
    x = fx.s1_iid(64, seed=9).repeat(40, 1, 1, 1, 1).cuda()          # 2560 >= 3 * 592
    sched(-1)
    auto_big, auto_small = layer(x), layer(x[:300])
    sched(0)
    assert torch.equal(layer(x), auto_big) and torch.equal(layer(x[:300]), auto_small)
    sched(1)
    assert torch.equal(layer(x), auto_big) and torch.equal(layer(x[:300]), auto_small)
