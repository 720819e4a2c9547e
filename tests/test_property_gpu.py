"""Property tests (hypothesis) of the fused forward on the GPU: arbitrary skeletons -- any number of joints, bodies and
bones, bones that share or repeat a source joint, the same bone listed twice, chains and stars -- arbitrary lengths down to
the T = 129 minimum, hops, wavelengths and radar positions, both memory layouts.  Every draw is checked against the CPU
oracle with the layer's tiered criterion, and the two schedules of the kernel must agree bit for bit."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import virtual_radar_oracle as vro

pytestmark = pytest.mark.gpu


@st.composite
def cases(draw):
    V = draw(st.integers(2, 34))
    M = draw(st.integers(1, 4))
    E = draw(st.integers(1, 40))
    kind = draw(st.sampled_from(["random", "star", "chain", "duplicates"]))
    if kind == "star":            # every bone starts at one joint: one bone group gets them all
        hub = draw(st.integers(0, V - 1))
        edges = [(hub, (hub + 1 + i) % V) for i in range(min(E, 30))]
    elif kind == "chain":
        edges = [(i % V, (i + 1) % V) for i in range(E)]
    elif kind == "duplicates":    # the same bone several times, and both directions
        a, b = draw(st.integers(0, V - 1)), draw(st.integers(0, V - 1))
        b = b if b != a else (a + 1) % V
        edges = [(a, b), (a, b), (b, a)] + [(draw(st.integers(0, V - 1)), draw(st.integers(0, V - 1))) for _ in range(max(E - 3, 0))]
    else:
        edges = [(draw(st.integers(0, V - 1)), draw(st.integers(0, V - 1))) for _ in range(E)]
    edges = [(s, d if d != s else (s + 1) % V) for s, d in edges]      # a zero-length bone is 0/0 in the reference
    T = draw(st.sampled_from([129, 130, 200, 257, 300, 301, 512, 777, 1500]))
    N = draw(st.integers(1, 5))
    hop = draw(st.sampled_from([8, 16, 16, 16, 37]))
    lam = draw(st.sampled_from([5e-4, 9e-4, 1e-3, 5e-3]))
    loc = draw(st.sampled_from([(0., 0., 0.), (0.5, -1.0, 2.0), (0., 0., 3.5)]))
    layout = draw(st.sampled_from(["planar", "coordinate_innermost"]))
    seed = draw(st.integers(0, 2 ** 16))
    return dict(V=V, M=M, edges=edges, T=T, N=N, hop=hop, lam=lam, loc=loc, layout=layout, seed=seed)


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(cases())
def test_arbitrary_skeletons_match_the_oracle(c):
    from skeleton_action_recognition_b200 import VirtualRadar, _cabi
    g = torch.Generator().manual_seed(c["seed"])
    shape = (c["N"], c["T"], c["V"], c["M"], 3) if c["layout"] == "coordinate_innermost" else (c["N"], 3, c["T"], c["V"], c["M"])
    x = torch.randn(*shape, generator=g) * 0.4
    if c["layout"] == "coordinate_innermost":
        x = x.permute(0, 4, 1, 2, 3)
    kw = dict(edges=c["edges"], wavelength=c["lam"], radar_location=list(c["loc"]), hop_length=c["hop"])
    layer = VirtualRadar(device="cuda:0", **kw).to("cuda:0")
    out = layer(x.cuda())
    ref = vro.forward(x, edges=c["edges"], wavelength=c["lam"], radar_location=c["loc"], hop_length=c["hop"],
                      distance=vro.distance_mode_for(x)).numpy()
    assert tuple(out.shape) == ref.shape == (c["N"], 256, c["T"] // c["hop"] + 1)
    rep = vro.parity_report(out.cpu().numpy(), ref)
    assert vro.parity_ok(rep), (c, rep)
    _cabi.set_schedule(1)
    try:
        assert torch.equal(layer(x.cuda()), out), c
    finally:
        _cabi.set_schedule(-1)
