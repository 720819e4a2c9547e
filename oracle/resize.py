"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the consumer's input stage that follows the layer.

Reference models/resnet.py:24-26:
    x = self.virtual_radar(x)                                  # (N, n_fft, F)
    x = x.unsqueeze(dim=1)                                     # (N, 1, n_fft, F)
    x = torch.nn.functional.interpolate(x, self.image_size)    # (N, 1, S, S), mode='nearest'

`interpolate(..., size=S)` with the default mode is ATen's legacy nearest neighbour: output index d
reads input index min(int(floorf(d * scale)), in - 1), scale = float(in) / out, evaluated in float32
(aten/src/ATen/native/UpSample.h, nearest_neighbor_compute_source_index / compute_scales_value).
Pinned in tests/test_oracle.py against torch.nn.functional.interpolate itself (torch is present on
both the build container and the GPU box) over a sweep of (in, out) sizes.
"""
import numpy as np


def nearest_index(out_size, in_size):
    """Source index of every output index, float32 arithmetic exactly as ATen does it."""
    scale = np.float32(in_size) / np.float32(out_size)
    d = np.arange(out_size, dtype=np.float32)
    idx = np.floor(d * scale).astype(np.int64)          # float32 product, then floor
    return np.minimum(idx, in_size - 1)


def resize_nearest(spec, image_size):
    """spec (N, R, F) -> (N, 1, S, S) as models/resnet.py:25-26 does."""
    spec = np.asarray(spec)
    rows = nearest_index(image_size, spec.shape[1])
    cols = nearest_index(image_size, spec.shape[2])
    return spec[:, rows][:, :, cols][:, None]


def kept_frames(n_frames, image_size):
    """The distinct STFT frames the resize reads (what a fused kernel has to transform)."""
    return np.unique(nearest_index(image_size, n_frames))
