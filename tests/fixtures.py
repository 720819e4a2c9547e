"""Shared builders for the parity inputs (SURVEY.md 8d synthetic inputs S1-S3 + golden crops)."""
import os

import numpy as np
import torch

from oracle.pad_frames import pad_frames, notebook_tensor
from oracle.virtual_radar_oracle import NTU_EDGES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def s1_iid(n, seed=0, shape=(3, 300, 25, 2), scale=0.3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, *shape, generator=g) * scale


def s2_ntu_like(n, seed=1):
    """S1 with the second body zeroed with p=0.6 and frames t>=L zeroed, L~U{50..300}."""
    x = s1_iid(n, seed=0)
    g = torch.Generator().manual_seed(seed)
    drop = torch.rand(n, generator=g) < 0.6
    length = torch.randint(50, 301, (n,), generator=g)
    x[drop, :, :, :, 1] = 0
    t = torch.arange(300)
    x = x * (t[None, :] < length[:, None])[:, None, :, None, None]
    return x


def s3_smooth(n, seed=2, T=300, V=25, M=2):
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(n, 3, 1, V, M, generator=g) * 0.3
    t = torch.arange(T, dtype=torch.float32)[None, None, :, None, None]
    x = pose.expand(n, 3, T, V, M).clone()
    for k in (1, 2, 3):
        a = torch.randn(n, 3, 1, V, M, generator=g) * 0.05
        ph = torch.rand(n, 3, 1, V, M, generator=g) * 2 * np.pi
        x = x + a * torch.sin(2 * np.pi * k * t / T + ph)
    return x


def golden_cases():
    """name -> (x tensor with the reference-side strides, kwargs, golden y, golden iq)."""
    out = {}
    z = load("ntu_raw.npz")
    out["ntu_raw"] = (torch.from_numpy(z["x"]), dict(wavelength=5e-4), z["y"], z["iq"])
    z = load("randn_small.npz")
    out["randn_small"] = (s1_iid(256)[:8], dict(wavelength=5e-4), z["y"], z["iq"])
    z = load("gait_crop.npz")
    out["gait_crop"] = (notebook_tensor(pad_frames(z["raw"], num_pad_frames=int(z["pad"]))),
                        dict(edges=[tuple(e) for e in z["edges"].tolist()], wavelength=5e-4), z["y"], z["iq"])
    z = load("cmu_crop.npz")
    out["cmu_crop"] = (notebook_tensor(pad_frames(z["raw"], num_pad_frames=int(z["pad"]))),
                       dict(edges=[tuple(e) for e in z["edges"].tolist()], wavelength=5e-3), z["y"], z["iq"])
    z = load("ntu_nb_crop.npz")
    out["ntu_nb_crop"] = (notebook_tensor(pad_frames(z["raw"], num_pad_frames=int(z["pad"]))),
                          dict(edges=NTU_EDGES, wavelength=9e-4), z["y"], z["iq"])
    z = load("offaxis.npz")
    g = torch.Generator().manual_seed(int(z["seed"]))
    x = torch.randn(*z["shape"].tolist(), generator=g) * float(z["scale"])
    out["offaxis"] = (x, dict(edges=[tuple(e) for e in z["edges"].tolist()], wavelength=1e-3,
                              radar_location=z["radar_location"].tolist()), z["y"], z["iq"])
    return out


GAIT_EDGES = [(0, 1), (1, 2), (1, 3), (3, 5), (5, 7), (1, 4), (4, 6), (6, 8), (0, 9),
              (9, 11), (11, 13), (13, 15), (0, 10), (10, 12), (12, 14), (14, 16)]

# BASELINE configs 1 and 3 at FULL size: notebook cell, known-answer row of BASELINE.md section 3, up-sampling factor
FULL = {"ntu": dict(row="B", pad=550, scale=1.0, kw=dict(edges=NTU_EDGES, wavelength=9e-4)),
        "cmu": dict(row="C", pad=20, scale=0.001, kw=dict(edges=[(i, i + 1) for i in range(41)], wavelength=5e-3)),
        "gait": dict(row="D", pad=10, scale=1.0, kw=dict(edges=GAIT_EDGES, wavelength=5e-4))}


def full_case(name):
    """The notebook's input of cells 4 / 2 / 3 rebuilt from the committed raw arrays (tests/golden/full_inputs.npz):
    pad_frames (reference utils.py:82-89, scipy float64) -> transpose(2,0,1) -> expand_dims -> torch.Tensor, i.e.
    (1,3,T,V,1) float32 with the coordinate axis innermost.  Returns (x, layer kwargs, golden dict)."""
    c = FULL[name]
    raw = load("full_inputs.npz")[name].astype(np.float64) if name == "cmu" else load("full_inputs.npz")[name]
    x = notebook_tensor(pad_frames(raw * c["scale"] if c["scale"] != 1.0 else raw, num_pad_frames=c["pad"]))
    g = load("full_outputs.npz")
    gold = {"y": g[name + "_y"], "iq": g[name + "_iq"], "rowsum": g[name + "_rowsum"], "stride": int(g["stride"]),
            "x_strides": tuple(int(s) for s in g[name + "_x_strides"]),
            "up": g[name + "_up"], "up_sum": float(g[name + "_up_sum"])}
    return x, c["kw"], gold


def full_raw(name):
    """The raw (T, V, 3) array of a full-size case as the notebook holds it before `pad_frames` (dtype as in the
    reference's files: NTU float32, CMU / gait float64; CMU scaled to metres), and the up-sampling factor."""
    c = FULL[name]
    raw = load("full_inputs.npz")[name]
    if name == "cmu":
        raw = raw.astype(np.float64) * c["scale"]
    return raw, c["pad"]
