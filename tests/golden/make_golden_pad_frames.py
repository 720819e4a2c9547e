"""Golden vectors for the temporal up-sampling pre-stage, made by running the REAL
reference code: /root/reference/utils.py is loaded verbatim by path and `Dataset.pad_frames`
(utils.py:134-140) and `Dataset.__getitem__`'s cast (utils.py:128-132) are executed on small inputs.

    python tests/golden/make_golden_pad_frames.py        (needs /root/reference; build container only)

utils.py imports plotting / data-generation modules that are absent here and irrelevant to pad_frames
(matplotlib, PIL, data_gen.gen_joint_data); they are satisfied by empty stub modules.  scipy and numpy,
which do the arithmetic, are the real packages.

Output: tests/golden/pad_frames_dataset.npz
  x_ntu   (2,3,300,25,2) f32  first two NTU example sequences         k=4   -> y_ntu  (2,3,1200,25,2) f32
  x_rand  (3,3,64,5,3)   f32  seeded randn (regenerated from the seed) k=250 -> y_rand (3,3,16000,5,3) f32 [every 37th frame kept]
  x_short (1,3,13,4,1)   f32  T = 13 = Gaussian radius + 1             k=9   -> y_short
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference_utils():
    for name in ("matplotlib", "matplotlib.pyplot", "PIL", "data_gen", "data_gen.gen_joint_data"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            m.__all__ = []
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_pad(mod, x, k, sigma=3):
    """Dataset.pad_frames + the FloatTensor cast of __getitem__, sample by sample, unbound (no files needed)."""
    out = []
    for s in x:
        me = types.SimpleNamespace(T=s.shape[-3], sigma=sigma, num_pad_frames=k)
        y = mod.Dataset.pad_frames(me, s)                       # utils.py:134-140  (float64)
        out.append(torch.from_numpy(y).type(torch.FloatTensor).numpy())   # utils.py:130-132
    return np.stack(out)


def main():
    mod = load_reference_utils()
    ntu = np.load(os.path.join(REF, "data", "NTU_preprocessed_skeleton_examples.npy"))[:2].astype(np.float32)
    y_ntu = ref_pad(mod, ntu, 4)
    g = torch.Generator().manual_seed(21)
    x_rand = (torch.randn(3, 3, 64, 5, 3, generator=g) * 0.4).numpy()
    y_rand = ref_pad(mod, x_rand, 250)[:, :, ::37]
    g = torch.Generator().manual_seed(22)
    x_short = (torch.randn(1, 3, 13, 4, 1, generator=g)).numpy()
    y_short = ref_pad(mod, x_short, 9)
    np.savez_compressed(os.path.join(HERE, "pad_frames_dataset.npz"), x_ntu=ntu, y_ntu=y_ntu, k_ntu=4,
                        y_rand=y_rand, k_rand=250, seed_rand=21, stride_rand=37,
                        x_short=x_short, y_short=y_short, k_short=9)
    print({k: v.shape for k, v in dict(y_ntu=y_ntu, y_rand=y_rand, y_short=y_short).items()})


if __name__ == "__main__":
    main()
