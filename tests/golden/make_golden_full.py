"""Full-size fixtures for BASELINE configs 1 and 3 (notebook cells 2-4), made by running the REAL reference here.

    python tests/golden/make_golden_full.py      (needs /root/reference; run in the build container)

Writes
  full_inputs.npz   the three RAW motion-capture arrays the notebook loads (data, not code): the NTU example
                    [0,:,:,:,0] (300,25,3 f32), cmu_mocap.npy (2751,42,3; stored as f32, which holds it exactly),
                    simulated_gait.npy (8192,17,3 f64).  The tests rebuild the notebook's inputs from them with
                    oracle.pad_frames (utils.py:82-89) + the notebook's transposes, so that T = 165 000 / 55 020 /
                    81 920 and the C-innermost strides are the notebook's.
  full_outputs.npz  from the real reference's forward (layers/virtual_radar.py loaded by path, nnAudio restated):
                    every STRIDE-th spectrogram column and every 16th baseband sample of each run (the full
                    outputs are 3.5-10.5 MB each), plus a float64 checksum per spectrogram row over ALL columns; and from
                    the real utils.pad_frames (utils.py:82-89): every 997th up-sampled frame (float32) and the sum.
known_answers.json (rows B, C, D: shape / sum / min / max / argmax) comes from make_golden.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import load_reference, ref_iq, GAIT_EDGES, REF  # noqa: E402
from make_golden_pad_frames import load_reference_utils  # noqa: E402
from oracle.pad_frames import pad_frames, notebook_tensor  # noqa: E402

STRIDE = 48


def main():
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    ntu = np.load(os.path.join(REF, "data", "NTU_preprocessed_skeleton_examples.npy"))
    nb = np.ascontiguousarray(np.transpose(ntu[0, :, :, :, 0], (1, 2, 0)))          # (300,25,3) f32
    cmu = np.load(os.path.join(REF, "data", "cmu_mocap.npy"))
    gait = np.load(os.path.join(REF, "data", "simulated_gait.npy"))
    assert np.array_equal(cmu.astype(np.float32).astype(np.float64), cmu)
    np.savez_compressed(os.path.join(HERE, "full_inputs.npz"), ntu=nb, cmu=cmu.astype(np.float32), gait=gait)

    cases = {
        "ntu": (nb, 550, dict(wavelength=9e-4)),                                                # cell 4 -> row B
        "cmu": (cmu * 0.001, 20, dict(edges=[(i, i + 1) for i in range(41)], wavelength=5e-3)),   # cell 2 -> row C
        "gait": (gait, 10, dict(edges=GAIT_EDGES, wavelength=5e-4)),                              # cell 3 -> row D
    }
    out = {"stride": STRIDE}
    ref_utils = load_reference_utils()                       # the REAL utils.py (pad_frames, utils.py:82-89)
    for name, (raw, pad, kw) in cases.items():
        up = ref_utils.pad_frames(raw, num_pad_frames=pad)
        assert np.array_equal(up, pad_frames(raw, num_pad_frames=pad))      # the oracle's restatement is the same function
        out[name + "_up"] = np.ascontiguousarray(up[::997].astype(np.float32))   # every 997th up-sampled frame, as torch.Tensor casts it
        out[name + "_up_sum"] = np.float64(up.astype(np.float32).astype(np.float64).sum())
        x = notebook_tensor(up)
        layer = ref.VirtualRadar(device="cpu", **kw)
        y, iq = ref_iq(layer, x)
        y, iq = y.numpy(), iq.numpy()
        out[name + "_y"] = np.ascontiguousarray(y[:, :, ::STRIDE])
        out[name + "_iq"] = np.ascontiguousarray(iq[:, ::16])
        out[name + "_rowsum"] = y.astype(np.float64).sum(axis=2)
        out[name + "_x_strides"] = np.array(x.stride())
        print(name, tuple(x.shape), tuple(x.stride()), y.shape, float(y.astype(np.float64).sum()), y.min(), y.max())
    np.savez_compressed(os.path.join(HERE, "full_outputs.npz"), **out)


if __name__ == "__main__":
    main()
