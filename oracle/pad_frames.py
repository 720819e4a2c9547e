"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's temporal up-sampling helpers.

  pad_frames            utils.py:82-89    (notebook variant: Gaussian smoothing along axis=1 of a
                                           (T,V,C) array, i.e. along the JOINT axis -- a quirk that
                                           is kept, SURVEY Appendix D -- then not-a-knot cubic
                                           interpolation in time to num_pad_frames*T frames)
  dataset_pad_frames    utils.py:134-140  (Dataset variant: smoothing along axis=-3 = time of a
                                           (3,T,V,M) sample)

Used to build the notebook-style inputs (configs 1 and 3) for parity tests.  float64 out, as in
the reference; the caller casts to float32 with torch.Tensor(...) exactly like the notebook.
"""
import numpy as np
from scipy.interpolate import interp1d
from scipy.ndimage import gaussian_filter1d


def pad_frames(data, num_pad_frames=1, sigma=3):
    frames = data.shape[0]
    smooth = gaussian_filter1d(data, sigma, axis=1)
    spline = interp1d(np.linspace(0, 1, frames), smooth, "cubic", axis=-3)
    return spline(np.linspace(0, 1, num_pad_frames * frames))


def dataset_pad_frames(sample, num_pad_frames=250, sigma=3):
    frames = sample.shape[-3]
    smooth = gaussian_filter1d(sample, sigma, axis=-3)
    spline = interp1d(np.linspace(0, 1, frames), smooth, "cubic", axis=-3)
    return spline(np.linspace(0, 1, num_pad_frames * frames))


def dataset_getitem(sample, num_pad_frames=250, sigma=3):
    """`Dataset.__getitem__` without the file access (utils.py:128-132): float32 sample (3,T,V,M) ->
    pad_frames (float64) -> FloatTensor.  This is what the GPU `pad_frames` must reproduce."""
    import torch
    return torch.from_numpy(dataset_pad_frames(sample, num_pad_frames, sigma)).type(torch.FloatTensor)


def notebook_tensor(data_tvc):
    """(T,V,C) float64 array -> the notebook's (1,3,T,V,1) float32 tensor with C-innermost
    strides (virtual_radar_example.ipynb cells 2-4: transpose(2,0,1), expand_dims, torch.Tensor)."""
    import torch
    a = data_tvc.transpose(2, 0, 1)
    a = np.expand_dims(a, axis=[0, -1])
    return torch.Tensor(a)
