"""Diagnostic (not a test): device pad_frames (Dataset variant) against scipy on several inputs: mismatching positions,
their magnitude relative to the coordinate range, and the kernel's time at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pad_frames as opf
from skeleton_action_recognition_b200 import pad_frames
tot = bad_tot = 0
for seed, shape, k in [(0, (2, 3, 300, 25, 2), 250), (1, (3, 3, 64, 5, 3), 250), (2, (2, 3, 300, 25, 2), 4), (3, (1, 3, 1000, 17, 1), 7),
                       (4, (2, 3, 40, 42, 1), 11), (5, (1, 3, 3000, 4, 1), 3), (6, (4, 3, 100, 25, 2), 97), (7, (2, 3, 300, 25, 2), 100)]:
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(*shape, generator=g) * 0.4).numpy()
    want = np.stack([opf.dataset_getitem(s, k).numpy() for s in x])
    got = pad_frames(torch.from_numpy(x).cuda(), k).cpu().numpy()
    bad = got != want
    tot += want.size; bad_tot += int(bad.sum())
    msg = ""
    if bad.any():
        ulp = np.abs(got[bad].view(np.int32).astype(np.int64) - want[bad].view(np.int32).astype(np.int64))
        msg = " max ulp %d, largest |value| among them %.3e (range %.2f)" % (ulp.max(), np.abs(want[bad]).max(), np.abs(want).max())
    print("seed %d shape %s k %d: %d of %d positions differ%s" % (seed, shape, k, bad.sum(), want.size, msg), flush=True)
print("total: %d of %d" % (bad_tot, tot))
x = (torch.randn(148, 3, 300, 25, 2) * 0.3).cuda(); o = torch.empty(148, 3, 75000, 25, 2, device="cuda")
for _ in range(2): pad_frames(x, 250, out=o)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): pad_frames(x, 250, out=o)
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
print("pad_frames N=148 k=250: %.3f ms, %.1f k seq/s, %.3f of 6541 GB/s" % (ms, 148 / ms, (o.numel() + x.numel()) * 4 / ms / 1e6 / 6541.1))
