"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the one third-party routine on the hot path.

`nnAudio.Spectrogram.STFT` (package nnAudio, unpinned in the reference's
requirements.txt:2; the notebook's pip log shows 0.1.1, and the `device=` kwarg the
reference passes at layers/virtual_radar.py:76 only exists in 0.1.0-0.1.5).  nnAudio's
source is NOT under /root/reference and cannot be installed (no network), so its published
algorithm is restated here from the package documentation / SURVEY.md Appendix B:

  * Fourier kernels  wsin[k,0,s] = w[s]*sin(2*pi*k*s/n_fft),  wcos likewise with cos,
    k = 0..freq_bins-1 (freq_scale='no'), w = scipy.signal.get_window('hann', n_fft,
    fftbins=True); computed in float64, stored as float32 parameters `wsin`, `wcos` of shape
    (freq_bins, 1, n_fft).
  * forward (center=True, pad_mode='reflect', output_format='Complex'):
    (N, L) -> (N, 1, L) -> ReflectionPad1d(n_fft//2) -> conv1d(., wsin, stride=hop) and
    conv1d(., wcos, stride=hop) -> stack((real, -imag), -1)  =>  (N, freq_bins, L//hop+1, 2).

Call sites in the reference: layers/virtual_radar.py:3 (import), :71-76 (construct),
:124-125 (two forwards, one on I and one on Q).

Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.
"""
import time

import numpy as np
import torch
from scipy.signal import get_window


def fourier_kernels(n_fft, freq_bins=None, window="hann"):
    """Windowed DFT kernels of nnAudio's create_fourier_kernels(freq_scale='no')."""
    if freq_bins is None:
        freq_bins = n_fft // 2 + 1
    s = np.arange(0, n_fft, 1.0)
    win = get_window(window, int(n_fft), fftbins=True)
    wsin = np.empty((freq_bins, 1, n_fft))
    wcos = np.empty((freq_bins, 1, n_fft))
    for k in range(freq_bins):
        wsin[k, 0, :] = win * np.sin(2 * np.pi * k * s / n_fft)
        wcos[k, 0, :] = win * np.cos(2 * np.pi * k * s / n_fft)
    return wsin.astype(np.float32), wcos.astype(np.float32)


class STFT(torch.nn.Module):
    """Drop-in for nnAudio 0.1.x `Spectrogram.STFT` restricted to what the reference uses."""

    def __init__(self, n_fft=2048, freq_bins=None, hop_length=512, window="hann",
                 freq_scale="no", center=True, pad_mode="reflect", fmin=50, fmax=6000,
                 sr=22050, trainable=False, output_format="Complex", device="cuda:0",
                 verbose=False):
        super().__init__()
        if freq_scale != "no" or not center or pad_mode != "reflect":
            raise NotImplementedError("only the configuration used by layers/virtual_radar.py:71-76")
        if output_format != "Complex":
            raise NotImplementedError("only output_format='Complex'")
        self.stride = hop_length
        self.n_fft = n_fft
        self.freq_bins = freq_bins
        self.output_format = output_format
        self.pad_amount = n_fft // 2
        start = time.time()
        wsin, wcos = fourier_kernels(n_fft, freq_bins=freq_bins, window=window)
        self.wsin = torch.nn.Parameter(torch.tensor(wsin, device=device), requires_grad=trainable)
        self.wcos = torch.nn.Parameter(torch.tensor(wcos, device=device), requires_grad=trainable)
        if verbose:
            print("STFT kernels created, time used = {:.4f} seconds".format(time.time() - start))

    def forward(self, x):
        if x.dim() == 2:
            x = x[:, None, :]
        elif x.dim() == 1:
            x = x[None, None, :]
        x = torch.nn.functional.pad(x, (self.pad_amount, self.pad_amount), mode="reflect")
        spec_imag = torch.nn.functional.conv1d(x, self.wsin, stride=self.stride)
        spec_real = torch.nn.functional.conv1d(x, self.wcos, stride=self.stride)
        spec_real = spec_real[:, :self.freq_bins, :]
        spec_imag = spec_imag[:, :self.freq_bins, :]
        return torch.stack((spec_real, -spec_imag), -1)
