"""TEST INFRASTRUCTURE ONLY -- gradients of the VirtualRadar layer with respect to `wavelength` and
`radar_location` (reference layers/virtual_radar.py:40-41, 65-69: trainable nn.Parameters; the reference
obtains the gradients from PyTorch autograd over forward(), :79-134).

Two independent routes, both on the CPU:
  autograd_grads   the oracle's own torch graph (oracle/virtual_radar_oracle.py, the reference's ops in the
                   reference's order) differentiated by torch.autograd, in float64 ("truth") or float32 (what
                   the reference itself would return).
  analytic_grads   the closed form the CUDA kernels implement (csrc/vr_backward.cuh), written with numpy in
                   float64: adjoint of the windowed STFT + log magnitude, then the derivatives of
                   z = sum amp * exp(j theta).  tests/test_oracle.py pins it to autograd_grads.
"""
import numpy as np
import torch

from . import virtual_radar_oracle as vro


def autograd_grads(x, grad_out, edges=vro.NTU_EDGES, wavelength=1e-3, radar_location=(0., 0., 0.),
                   n_fft=256, hop_length=16, dtype=torch.float64, wrt_x=False):
    """-> (dL/dwavelength, dL/dradar_location, out) or, with wrt_x, (.., .., dL/dx)."""
    o = vro.OracleVirtualRadar(edges, wavelength, radar_location, n_fft, hop_length, dtype)
    lam = o.wavelength.clone().requires_grad_(True)
    loc = o.radar_location.clone().requires_grad_(True)
    xx = x.to(dtype).clone().requires_grad_(wrt_x)
    iq = vro.synthesize_iq(xx, o.src, o.dst, loc, lam, "aten")
    out = vro.stft_logmag(iq, o.stft, n_fft)
    loss = (out * torch.as_tensor(grad_out).to(dtype)).sum()
    grads = torch.autograd.grad(loss, (lam, loc, xx) if wrt_x else (lam, loc))
    if wrt_x:
        return float(grads[0]), grads[1].numpy().astype(np.float64), grads[2].numpy().astype(np.float64)
    return float(grads[0]), grads[1].numpy().astype(np.float64), out.detach().numpy()


def _reflect(t, T):
    t = np.abs(t)
    return np.where(t >= T, 2 * (T - 1) - t, t)


def stft_adjoint(iq, grad_out, n_fft=256, hop=16):
    """iq (N,T,2), grad_out (N,n_fft,F) -> dL/d(iq) (N,T,2), float64."""
    iq = np.asarray(iq, np.float64)
    g = np.asarray(grad_out, np.float64)
    N, T, _ = iq.shape
    F = T // hop + 1
    z = iq[..., 0] + 1j * iq[..., 1]
    n = np.arange(n_fft)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)
    gz = np.zeros((N, T), np.complex128)
    for f in range(F):
        idx = _reflect(f * hop - n_fft // 2 + n, T)
        X = np.fft.fft(z[:, idx] * w, axis=1)                     # X[k] = sum w zp e^{-j 2 pi k n / n_fft}
        a = np.abs(X)
        gk = np.roll(g[:, :, f], -(n_fft // 2), axis=1)           # out row r shows bin (r - n_fft/2) mod n_fft
        G = np.where(a > 0, gk * X / (a * (a + 1e-6) + (a == 0)), 0)
        y = w * (np.fft.ifft(G, axis=1) * n_fft)                  # sum_k G e^{+j 2 pi k n / n_fft}
        np.add.at(gz, (slice(None), idx), y)
    return np.stack((gz.real, gz.imag), -1)


def analytic_grads(x, grad_out, edges=vro.NTU_EDGES, wavelength=1e-3, radar_location=(0., 0., 0.),
                   n_fft=256, hop_length=16, wrt_x=False):
    x64 = np.asarray(x, np.float64)
    lam = float(np.float32(wavelength))
    L = np.asarray(np.float32(radar_location), np.float64)[None, :, None, None, None]
    src, dst = map(list, zip(*edges))
    S, D = x64[:, :, :, src], x64[:, :, :, dst]                  # (N,3,T,E,M)
    rng = np.sqrt(((S - L) ** 2).sum(1))
    th = 4 * np.pi * rng / lam
    A, B = L - (S + D) / 2, D - S
    na, nb = np.sqrt((A ** 2).sum(1)), np.sqrt((B ** 2).sum(1))
    dot = (A * B).sum(1)
    q = na * nb + 1e-6
    u = dot / q
    cbar = nb.mean(axis=2, keepdims=True)
    c = cbar ** 2
    K = np.sqrt(np.pi) * cbar
    den = 1 + (c - 1) * u ** 2
    amp = K / den
    iq = np.stack(((amp * np.cos(th)).sum((2, 3)), (amp * np.sin(th)).sum((2, 3))), -1)
    gz = stft_adjoint(iq, grad_out, n_fft, hop_length)
    gI, gQ = gz[..., 0][:, :, None, None], gz[..., 1][:, :, None, None]
    dth = amp * (gQ * np.cos(th) - gI * np.sin(th))
    damp = gI * np.cos(th) + gQ * np.sin(th)
    g_lam = (dth * (-th / lam)).sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        kth = np.where(rng > 0, dth * (4 * np.pi / lam) / rng, 0)[:, None]
        kamp = (damp * (-K * 2 * u * (c - 1) / den ** 2))[:, None]
        dudA = np.where(na[:, None] > 0, B / q[:, None] - (dot * nb / (na * q ** 2))[:, None] * A, 0)
    g_loc = (kth * (L - S) + kamp * dudA).sum((0, 2, 3, 4))
    if not wrt_x:
        return float(g_lam), g_loc, iq
    # dL/dx: through the range (S), the aspect cosine (A = L - (S+D)/2, B = D - S) and the mean bone length
    with np.errstate(divide="ignore", invalid="ignore"):
        dudB = np.where(nb[:, None] > 0, A / q[:, None] - (dot * na / (nb * q ** 2))[:, None] * B, 0)
        unitB = np.where(nb[:, None] > 0, B / nb[:, None], 0)
    G_c = (damp * np.sqrt(np.pi) * (1 / den - 2 * c * u ** 2 / den ** 2)).sum(2, keepdims=True)     # (N,T,1,M)
    gB = kamp * dudB + (G_c / len(src))[:, None] * unitB
    gS = kth * (S - L) - 0.5 * kamp * dudA - gB
    gD = -0.5 * kamp * dudA + gB
    gx = np.zeros_like(x64)
    for e, (si, di) in enumerate(zip(src, dst)):
        gx[:, :, :, si] += gS[:, :, :, e]
        gx[:, :, :, di] += gD[:, :, :, e]
    return float(g_lam), g_loc, gx
