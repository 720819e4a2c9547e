"""Generate the committed golden fixtures by running the REAL reference in this container.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container)

The reference file /root/reference/layers/virtual_radar.py is loaded verbatim BY PATH (nothing is
copied into the repo).  Its one missing import, nnAudio.Spectrogram.STFT (third-party, absent, not
installable), is satisfied by oracle/nnaudio_stft.py injected into sys.modules -- the restatement
described in SURVEY.md Appendix B.  Everything else (geometry, RCS, synthesis, combination, log,
roll) is the reference's own code executing on CPU float32.

Outputs (tests/golden/*.npz, small):
  ntu_raw.npz      x = first 2 samples of data/NTU_preprocessed_skeleton_examples.npy (2,3,300,25,2),
                   lambda=5e-4 -> y (2,256,19), plus intermediate I/Q (2,300,2)
  randn_small.npz  seeded randn*0.3 (8,3,300,25,2), lambda=5e-4 -> y, I/Q   (input regenerated from the seed)
  gait_crop.npz    data/simulated_gait.npy[:640] (f64, 17 joints), notebook cell 3 recipe (pad x10,
                   16 edges, lambda=5e-4, C-innermost strides) -> y (1,256,401)
  cmu_crop.npz     data/cmu_mocap.npy[:400]*0.001 (42 joints), cell 2 recipe (pad x20, chain edges,
                   lambda=5e-3) -> y (1,256,501)
  ntu_nb_crop.npz  NTU example [0,:,:,:,0] first 40 frames, cell 4 recipe (pad x100, lambda=9e-4)
  offaxis.npz      seeded randn, radar at (0.5,-1.0,2.0), lambda=1e-3, M=3, V=10, 9 edges, T=200
  known_answers.json  shape/sum/min/max/argmax of the full-size runs A-E of BASELINE.md section 3
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import nnaudio_stft  # noqa: E402
from oracle.pad_frames import pad_frames, notebook_tensor  # noqa: E402


def load_reference():
    pkg = types.ModuleType("nnAudio")
    sub = types.ModuleType("nnAudio.Spectrogram")
    sub.STFT = nnaudio_stft.STFT
    pkg.Spectrogram = sub
    sys.modules["nnAudio"] = pkg
    sys.modules["nnAudio.Spectrogram"] = sub
    spec = importlib.util.spec_from_file_location("ref_virtual_radar",
                                                  os.path.join(REF, "layers", "virtual_radar.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


GAIT_EDGES = [(0, 1), (1, 2), (1, 3), (3, 5), (5, 7), (1, 4), (4, 6), (6, 8), (0, 9),
              (9, 11), (11, 13), (13, 15), (0, 10), (10, 12), (12, 14), (14, 16)]


def ref_iq(layer, x):
    """Intermediate I/Q of the reference: re-run its forward up to the sum by hooking the STFT."""
    grabbed = []
    orig = layer.stft.forward

    def spy(sig):
        grabbed.append(sig.detach().clone())
        return orig(sig)
    layer.stft.forward = spy
    with torch.no_grad():
        y = layer(x)
    layer.stft.forward = orig
    return y, torch.stack(grabbed[:2], dim=-1)


def stats(y):
    y = y.numpy()
    am = np.unravel_index(np.argmax(y), y.shape)
    return {"shape": list(y.shape), "sum": float(y.astype(np.float64).sum()),
            "min": float(y.min()), "max": float(y.max()), "argmax": [int(i) for i in am],
            "y_0_128_0": float(y[0, 128, 0]), "y_0_0_last": float(y[0, 0, -1])}


def main():
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    ka = {}

    ntu = np.load(os.path.join(REF, "data", "NTU_preprocessed_skeleton_examples.npy"))
    # --- A: raw NTU batch -------------------------------------------------------------------
    layer = ref.VirtualRadar(wavelength=5e-4, device="cpu")
    xa = torch.from_numpy(ntu.copy())
    ya, iqa = ref_iq(layer, xa)
    ka["A"] = stats(ya)
    np.savez_compressed(os.path.join(HERE, "ntu_raw.npz"), x=ntu[:2], y=ya[:2].numpy(), iq=iqa[:2].numpy(),
                        wavelength=5e-4)

    # --- E: seeded randn (full size known-answer) + small committed slice ----------------------
    g = torch.Generator().manual_seed(0)
    xe = torch.randn(256, 3, 300, 25, 2, generator=g) * 0.3
    with torch.no_grad():
        ye = layer(xe)
    ka["E"] = stats(ye)
    ys, iqs = ref_iq(layer, xe[:8])
    assert torch.equal(ys, ye[:8])
    np.savez_compressed(os.path.join(HERE, "randn_small.npz"), y=ys.numpy(), iq=iqs.numpy(), seed=0, n=8,
                        wavelength=5e-4)

    # --- D / gait (cell 3) ------------------------------------------------------------------
    gait = np.load(os.path.join(REF, "data", "simulated_gait.npy"))
    lay = ref.VirtualRadar(edges=GAIT_EDGES, wavelength=5e-4, device="cpu")
    with torch.no_grad():
        ka["D"] = stats(lay(notebook_tensor(pad_frames(gait, num_pad_frames=10))))
    crop = gait[:640].copy()
    xg = notebook_tensor(pad_frames(crop, num_pad_frames=10))
    yg, iqg = ref_iq(lay, xg)
    np.savez_compressed(os.path.join(HERE, "gait_crop.npz"), raw=crop, y=yg.numpy(), iq=iqg.numpy(),
                        edges=np.array(GAIT_EDGES), pad=10, wavelength=5e-4, x_strides=np.array(xg.stride()))

    # --- C / cmu (cell 2) -------------------------------------------------------------------
    cmu = np.load(os.path.join(REF, "data", "cmu_mocap.npy")) * 0.001
    cmu_edges = [(i, i + 1) for i in range(41)]
    lay = ref.VirtualRadar(edges=cmu_edges, wavelength=5e-3, device="cpu")
    with torch.no_grad():
        ka["C"] = stats(lay(notebook_tensor(pad_frames(cmu, num_pad_frames=20))))
    crop = cmu[:400].copy()
    xc = notebook_tensor(pad_frames(crop, num_pad_frames=20))
    yc, iqc = ref_iq(lay, xc)
    np.savez_compressed(os.path.join(HERE, "cmu_crop.npz"), raw=crop, y=yc.numpy(), iq=iqc.numpy(),
                        edges=np.array(cmu_edges), pad=20, wavelength=5e-3)

    # --- B / NTU notebook (cell 4) ------------------------------------------------------------
    nb = np.transpose(ntu[0, :, :, :, 0], (1, 2, 0))
    lay = ref.VirtualRadar(wavelength=9e-4, device="cpu")
    xb = notebook_tensor(pad_frames(nb, num_pad_frames=550))
    ka["B_input_sum"] = float(xb.double().sum())
    with torch.no_grad():
        ka["B"] = stats(lay(xb))
    crop = nb[:40].copy()
    xb = notebook_tensor(pad_frames(crop, num_pad_frames=100))
    yb, iqb = ref_iq(lay, xb)
    np.savez_compressed(os.path.join(HERE, "ntu_nb_crop.npz"), raw=crop, y=yb.numpy(), iq=iqb.numpy(),
                        pad=100, wavelength=9e-4)

    # --- off-axis radar, odd shapes -----------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    xo = torch.randn(3, 3, 200, 10, 3, generator=g) * 0.4
    eo = [(0, 1), (1, 2), (2, 3), (3, 4), (1, 5), (5, 6), (1, 7), (7, 8), (8, 9)]
    lay = ref.VirtualRadar(edges=eo, wavelength=1e-3, radar_location=[0.5, -1.0, 2.0], device="cpu")
    yo, iqo = ref_iq(lay, xo)
    np.savez_compressed(os.path.join(HERE, "offaxis.npz"), y=yo.numpy(), iq=iqo.numpy(), seed=7,
                        shape=np.array(xo.shape), scale=0.4, edges=np.array(eo), wavelength=1e-3,
                        radar_location=np.array([0.5, -1.0, 2.0]))

    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print(json.dumps(ka, indent=1))


if __name__ == "__main__":
    main()
