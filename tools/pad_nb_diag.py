"""Diagnostic (not a test): where does the device `utils.pad_frames` differ from scipy, and are those float32 rounding ties?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import fixtures as fx
from oracle.pad_frames import pad_frames as host_pad
from skeleton_action_recognition_b200 import pad_frames_notebook
for name in sorted(fx.FULL):
    raw, k = fx.full_raw(name)
    up64 = host_pad(raw, k)                                   # (kT, V, C) float64 (scipy)
    want = up64.astype(np.float32)
    got = pad_frames_notebook(torch.from_numpy(raw).cuda(), k)[0, :, :, :, 0].permute(1, 2, 0).cpu().numpy()
    bad = np.argwhere(got != want)
    print(name, raw.dtype, "mismatches", len(bad), "of", want.size, flush=True)
    for idx in bad[:12]:
        i, v, c = idx
        w64 = up64[i, v, c]; a, b = want[i, v, c], got[i, v, c]
        lo, hi = (a, b) if a < b else (b, a)
        mid = (np.float64(lo) + np.float64(hi)) / 2
        ulp = abs(np.float64(hi) - np.float64(lo))
        print("  frame %d joint %d coord %d: scipy64 %.17g -> f32 %.9g | device %.9g | distance of scipy64 from the midpoint / ulp32 = %.3e | input frame pos %.6f"
              % (i, v, c, w64, a, b, abs(w64 - mid) / ulp, i * (raw.shape[0] - 1) / (k * raw.shape[0] - 1)))
    if len(bad):
        fr = bad[:, 0]
        pos = fr * (raw.shape[0] - 1) / (k * raw.shape[0] - 1)
        print("  frames: min %d max %d; fractional positions in interval: min %.4f max %.4f; exact grid points: %d" % (fr.min(), fr.max(), (pos % 1).min(), (pos % 1).max(), int(((pos % 1) == 0).sum())))
