"""Input side of the hot path: the reference's `utils.Dataset` (utils.py:92-140) with the
temporal up-sampling moved off the CPU workers.

The reference's `__getitem__` smooths and up-samples every sample on a DataLoader worker (scipy, 0.16 s per
sample) and hands a `num_pad_frames`-times larger float32 tensor (45 MB at the default 250) to the H2D copy.
Here `__getitem__` returns the RAW (3, T, V, M) sample straight from the memory-mapped `.npy`
(data_gen/gen_joint_data.py writes it) and `gpu_batches` copies pinned raw batches to the device and runs
`pad_frames` there (C ABI `vr_pad_frames_f32`): same constructor, same labels, same values after the
up-sampling, 250x less host work and PCIe traffic.
"""
import pickle
from pathlib import Path

import numpy as np
import torch

from .upsample import pad_frames


class Dataset(torch.utils.data.Dataset):
    """Same constructor as the reference (`utils.py:105`): `data_path` is the `(N,3,T,V,M)` float32 `.npy`,
    `label_path` the pickle holding `(sample_names, labels)`.  Samples come back raw; call
    `upsample(batch)` on a CUDA batch (or iterate `gpu_batches`) for what the reference's
    `__getitem__` would have produced."""

    def __init__(self, data_path, label_path, num_pad_frames=250, sigma=3):
        self.sigma = sigma
        self.num_pad_frames = num_pad_frames
        label_path, data_path = Path(label_path), Path(data_path)
        if not label_path.exists():
            raise FileNotFoundError("label file %s does not exist" % label_path)
        if not data_path.exists():
            raise FileNotFoundError("data file %s does not exist" % data_path)
        with open(label_path, "rb") as f:
            _, labels = pickle.load(f, encoding="latin1")
        self.data = np.load(data_path, allow_pickle=True, mmap_mode="r")
        self.labels = np.array(labels)
        if self.data.ndim != 5 or self.data.shape[1] != 3:
            raise ValueError("expected data of shape (N,3,T,V,M), got %s" % (self.data.shape,))
        if len(self.labels) != len(self.data):
            raise ValueError("%d labels for %d samples" % (len(self.labels), len(self.data)))
        self.T = self.data.shape[-3]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        x = torch.from_numpy(np.array(self.data[index], dtype=np.float32))     # a writable copy of the memory-mapped sample
        return x, torch.as_tensor(self.labels[index])

    def upsample(self, batch):
        """(N,3,T,V,M) or (3,T,V,M) CUDA float32 -> `num_pad_frames * T` frames, as the reference's
        `Dataset.pad_frames` + FloatTensor cast would give (utils.py:128-140)."""
        return pad_frames(batch, self.num_pad_frames, self.sigma)


def gpu_batches(loader, device, upsample=True, num_pad_frames=None, sigma=None, prefetch=True):
    """Iterate a DataLoader over `Dataset`: pinned host batch -> asynchronous H2D copy of the RAW samples ->
    up-sampling on the device.  Yields `(x_cuda, labels_cuda)`.  With `upsample=False` the raw batch is
    yielded (for `Model(num_pad_frames=...)`, which up-samples inside its forward).

    `prefetch=True` (default): double buffering -- the copy of batch k+1 is issued on a side stream (from a pinned staging
    copy) while the consumer works on batch k on the current stream; the yielded tensors are made safe to use on the
    current stream with `wait_stream` / `record_stream`, so the caller needs no extra synchronisation."""
    device = torch.device(device)
    ds = loader.dataset
    k = ds.num_pad_frames if num_pad_frames is None else num_pad_frames
    s = ds.sigma if sigma is None else sigma

    def finish(x, y):
        return (pad_frames(x, k, s) if upsample else x), y

    if not prefetch:
        for x, y in loader:
            if not x.is_pinned():
                x = x.pin_memory()
            yield finish(x.to(device, non_blocking=True), y.to(device, non_blocking=True))
        return

    copy_stream = torch.cuda.Stream(device)

    def stage(batch):
        x, y = batch
        if not x.is_pinned():
            x = x.pin_memory()
        with torch.cuda.stream(copy_stream):
            return x.to(device, non_blocking=True), y.to(device, non_blocking=True), x      # keep the pinned source alive

    it = iter(loader)
    try:
        nxt = stage(next(it))
    except StopIteration:
        return
    while nxt is not None:
        xd, yd, _pinned = nxt
        try:
            nxt = stage(next(it))                      # batch k+1 starts copying before batch k is handed out
        except StopIteration:
            nxt = None
        cur = torch.cuda.current_stream(device)
        cur.wait_stream(copy_stream)                   # batch k's copy is done before anything on the current stream reads it
        xd.record_stream(cur)
        yd.record_stream(cur)
        yield finish(xd, yd)
