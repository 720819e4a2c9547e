"""Drop-in replacement for the reference's `layers/virtual_radar.py` (class `VirtualRadar`,
module-level `edges`): same constructor arguments, same `forward(x)` signature, same
`state_dict` keys -- but `forward` is one fused sm_100a CUDA kernel reached through the C ABI of
include/virtual_radar_b200.h.  PyTorch is used only for device memory and streams.

Reference behaviour mirrored here (file:line in /root/reference):
  * constructor kwargs and defaults                     layers/virtual_radar.py:36-45
  * parameters `wavelength` (0-d), `radar_location` (3,) layers/virtual_radar.py:65-69
  * `self.src`, `self.dst` python lists                  layers/virtual_radar.py:70
  * `self.stft` with parameters wsin/wcos (n_fft,1,n_fft) layers/virtual_radar.py:71-76 (nnAudio)
  * forward: (N,3,T,V,M) f32 -> (N, n_fft, T//hop+1) f32  layers/virtual_radar.py:79-134

There is no CPU path and no PyTorch fallback: a CPU tensor, a missing shared object or a shape the
kernels do not cover raises.
"""
import ctypes
import math

import numpy as np
import torch

from .. import _cabi

# Default skeleton: the reference's NTU RGB+D bone list (layers/virtual_radar.py:10-13),
# grouped here by limb.  24 bones, 25 joints, 18 distinct source joints.
_SPINE = [(0, 1), (1, 20), (20, 2), (2, 3)]
_LEFT_ARM = [(20, 4), (4, 5), (5, 6), (6, 7), (7, 21), (7, 22)]
_RIGHT_ARM = [(20, 8), (8, 9), (9, 10), (10, 11), (11, 23), (11, 24)]
_HIPS = [(0, 16), (0, 12)]
_LEFT_LEG = [(12, 13), (13, 14), (14, 15)]
_RIGHT_LEG = [(16, 17), (17, 18), (18, 19)]
edges = _SPINE + _LEFT_ARM + _RIGHT_ARM + _HIPS + _LEFT_LEG + _RIGHT_LEG


class _STFTKernels(torch.nn.Module):
    """Holds `wsin` / `wcos` so that `state_dict()` has the reference's `stft.wsin`, `stft.wcos`
    entries (nnAudio 0.1.x STFT parameters).  The CUDA path evaluates the
    same windowed DFT with an FFT and does not read these tensors; `assert_dft()` verifies that a
    loaded checkpoint still holds the analytic Hann-windowed Fourier kernels."""

    def __init__(self, n_fft, hop_length, trainable, device):
        super().__init__()
        self.n_fft, self.stride = n_fft, hop_length
        wsin, wcos = self.analytic(n_fft)
        self.wsin = torch.nn.Parameter(wsin.to(device), requires_grad=trainable)
        self.wcos = torch.nn.Parameter(wcos.to(device), requires_grad=trainable)

    @staticmethod
    def analytic(n_fft):
        s = np.arange(n_fft, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2 * np.pi * s / n_fft)          # scipy get_window('hann', fftbins=True)
        ang = 2 * np.pi * np.arange(n_fft, dtype=np.float64)[:, None] * s[None, :] / n_fft
        wsin = (window * np.sin(ang)).astype(np.float32)[:, None, :]
        wcos = (window * np.cos(ang)).astype(np.float32)[:, None, :]
        return torch.from_numpy(wsin), torch.from_numpy(wcos)

    def is_dft(self):
        wsin, wcos = self.analytic(self.n_fft)
        return bool(torch.allclose(self.wsin.detach().cpu(), wsin, atol=1e-6)
                    and torch.allclose(self.wcos.detach().cpu(), wcos, atol=1e-6))

    def assert_dft(self):
        if not self.is_dft():
            raise NotImplementedError("stft.wsin/wcos differ from the Hann-windowed Fourier kernels")

    def logmag(self, iq):
        """General-kernel path (trained or trainable `wsin`/`wcos`, or an `n_fft` the fused FFT does not cover): what the
        reference computes at layers/virtual_radar.py:124-133 with nnAudio's conv1d STFT, as one float32-accurate GEMM per
        batch on the tensor cores (C ABI vr_stft_general_f32: tcgen05.mma kind::tf32, 3-term split, accumulator in tensor
        memory, magnitude / log / roll in the epilogue), differentiable with respect to `wsin`, `wcos` and `iq`
        (vr_stft_general_backward_f32).  iq (N,T,2) CUDA float32 -> (N, n_fft, T//hop+1)."""
        if not iq.is_cuda:
            raise RuntimeError("the general-kernel STFT (B200) has no CPU path")
        return _GeneralSTFTFunction.apply(iq, self.wsin, self.wcos, self.n_fft, self.stride, torch.is_grad_enabled())

    def _logmag_torch(self, iq):
        """TEST CROSS-CHECK ONLY (never called by forward): the same computation with torch ops (cuBLAS GEMM + autograd)."""
        n = self.n_fft
        zp = torch.nn.functional.pad(iq.permute(0, 2, 1), (n // 2, n // 2), mode="reflect")       # (N,2,T+n)
        frames = zp.unfold(2, n, self.stride)                                                       # (N,2,F,n)
        wc, ws = self.wcos[:, 0, :].t(), self.wsin[:, 0, :].t()                                     # (n, n_fft)
        c, s_ = torch.matmul(frames, wc), torch.matmul(frames, ws)                                  # (N,2,F,n_fft)
        real = c[:, 0] + s_[:, 1]            # stft(I).real - stft(Q).imag, nnAudio's imag = -conv(x, wsin)
        imag = c[:, 1] - s_[:, 0]            # stft(I).imag + stft(Q).real
        mag = torch.sqrt(real * real + imag * imag)
        out = torch.log(mag + 1e-6).transpose(1, 2)                                                 # (N,n_fft,F)
        return torch.roll(out, n // 2, dims=1)


class _GeneralSTFTFunction(torch.autograd.Function):
    """STFT against general kernels + log-magnitude + roll: forward and backward are the tcgen05 GEMM kernels of
    csrc/vr_stft_gemm.cuh behind vr_stft_general_f32 / vr_stft_general_backward_f32 (no library GEMM, no autograd graph)."""

    @staticmethod
    def forward(ctx, iq, wsin, wcos, n_fft, hop, grad_mode):
        iq = iq.contiguous()
        if iq.dtype != torch.float32 or wsin.dtype != torch.float32:
            raise ValueError("the general-kernel STFT computes in float32")
        N, T, _ = iq.shape
        L = _cabi.lib()
        parts = (ctypes.c_int64 * 3)()
        L.vr_stft_general_workspace_floats(N, T, n_fft, hop, parts)
        dev = iq.device
        frames = torch.empty(max(int(parts[0]), 1), dtype=torch.float32, device=dev)
        bt = torch.empty(max(int(parts[1]), 1), dtype=torch.float32, device=dev)
        # needs_input_grad ignores torch.no_grad(): the caller passes the grad mode, so that inference does not write Re / Im
        need = bool(grad_mode) and any(ctx.needs_input_grad[:3])
        csave = torch.empty(max(int(parts[2]), 1), dtype=torch.float32, device=dev) if need else None
        out = torch.empty((N, n_fft, T // hop + 1), dtype=torch.float32, device=dev)
        if N > 0:
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                rc = L.vr_stft_general_f32(iq.data_ptr(), N, T, n_fft, hop, wsin.contiguous().data_ptr(), wcos.contiguous().data_ptr(),
                                           frames.data_ptr(), bt.data_ptr(), csave.data_ptr() if need else None,
                                           out.data_ptr(), stream)
            _cabi.check(rc)
        if need:
            ctx.save_for_backward(frames, bt, csave)
        ctx.dims = (N, T, n_fft, hop, tuple(wsin.shape))
        return out

    @staticmethod
    def backward(ctx, gout):
        frames, bt, csave = ctx.saved_tensors
        N, T, n_fft, hop, wshape = ctx.dims
        dev = gout.device
        need_iq, need_w = ctx.needs_input_grad[0], (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        g = gout.contiguous().to(torch.float32)
        dc = torch.empty_like(csave)
        da = torch.empty(N * (T // hop + 1) * 2 * n_fft, dtype=torch.float32, device=dev) if need_iq else None   # frame gradients
        dbt = torch.empty_like(bt) if need_w else None
        giq = torch.empty((N, T, 2), dtype=torch.float32, device=dev) if need_iq else None
        gsin = torch.empty(wshape, dtype=torch.float32, device=dev) if need_w else None
        gcos = torch.empty(wshape, dtype=torch.float32, device=dev) if need_w else None
        if N > 0:
            ptr = lambda t: t.data_ptr() if t is not None else None     # noqa: E731
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                rc = _cabi.lib().vr_stft_general_backward_f32(g.data_ptr(), frames.data_ptr(), bt.data_ptr(), csave.data_ptr(),
                                                              N, T, n_fft, hop, dc.data_ptr(), ptr(da), ptr(dbt),
                                                              ptr(giq), ptr(gsin), ptr(gcos), stream)
            _cabi.check(rc)
        return (giq, gsin if ctx.needs_input_grad[1] else None, gcos if ctx.needs_input_grad[2] else None, None, None, None)


class _RadarFunction(torch.autograd.Function):
    """forward() with gradients for `wavelength`, `radar_location` and `x` (the reference gets them from
    PyTorch autograd over layers/virtual_radar.py:79-134; here: C ABI vr_backward_f32).  The
    forward launch is the same fused kernel, asked to also save the complex baseband signal."""

    @staticmethod
    def forward(ctx, lam, loc, xc, layer, flags):
        out, iq = layer._launch(xc, flags, want_iq=True)
        ctx.save_for_backward(lam, loc, xc, iq)
        ctx.layer, ctx.flags = layer, flags
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lam, loc, xc, iq = ctx.saved_tensors
        layer = ctx.layer
        N, _, T, V, M = xc.shape
        g = grad_out.contiguous().to(torch.float32)
        gz = torch.empty((N, T, 2), dtype=torch.float32, device=xc.device)
        gp = torch.zeros(4, dtype=torch.float64, device=xc.device)
        gx = torch.empty_like(xc) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(xc.device):
            stream = torch.cuda.current_stream(xc.device).cuda_stream
            rc = _cabi.lib().vr_backward_f32(xc.data_ptr(), iq.data_ptr(), g.data_ptr(), N, T, V, M,
                                             layer._src_c, layer._dst_c, len(layer.src),
                                             lam.data_ptr(), loc.data_ptr(), layer.n_fft, layer.hop_length,
                                             ctx.flags, gz.data_ptr(), gp.data_ptr(),
                                             gx.data_ptr() if gx is not None else None, ctypes.c_void_p(stream))
        _cabi.check(rc)
        g_lam = gp[0].to(torch.float32).reshape(lam.shape) if ctx.needs_input_grad[0] else None
        g_loc = gp[1:4].to(torch.float32).reshape(loc.shape) if ctx.needs_input_grad[1] else None
        return g_lam, g_loc, gx, None, None


class _SynthFunction(torch.autograd.Function):
    """Complex baseband synthesis only (layers/virtual_radar.py:93-123) -> iq (N,T,2), with gradients for
    `wavelength`, `radar_location` and `x` (C ABI vr_synth_adjoint_f32).  Used when the STFT is a general,
    trainable kernel pair and therefore runs outside the fused kernel."""

    @staticmethod
    def forward(ctx, lam, loc, xc, layer, flags):
        _, iq = layer._launch(xc, flags, want_iq=True)
        ctx.save_for_backward(lam, loc, xc)
        ctx.layer, ctx.flags = layer, flags
        return iq

    @staticmethod
    def backward(ctx, grad_iq):
        lam, loc, xc = ctx.saved_tensors
        layer = ctx.layer
        N, _, T, V, M = xc.shape
        g = grad_iq.contiguous().to(torch.float32)
        gp = torch.zeros(4, dtype=torch.float64, device=xc.device)
        gx = torch.empty_like(xc) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(xc.device):
            stream = torch.cuda.current_stream(xc.device).cuda_stream
            rc = _cabi.lib().vr_synth_adjoint_f32(xc.data_ptr(), g.data_ptr(), N, T, V, M, layer._src_c, layer._dst_c,
                                                  len(layer.src), lam.data_ptr(), loc.data_ptr(), ctx.flags, gp.data_ptr(),
                                                  gx.data_ptr() if gx is not None else None, ctypes.c_void_p(stream))
        _cabi.check(rc)
        g_lam = gp[0].to(torch.float32).reshape(lam.shape) if ctx.needs_input_grad[0] else None
        g_loc = gp[1:4].to(torch.float32).reshape(loc.shape) if ctx.needs_input_grad[1] else None
        return g_lam, g_loc, gx, None, None


class VirtualRadar(torch.nn.Module):
    """Skeleton sequences -> micro-Doppler log-spectrograms on a B200 (see module docstring)."""

    def __init__(self, edges=edges, wavelength=1e-3, radar_location=[0., 0., 0.],
                 train_wavelength=False, train_radar_location=False, train_stft_kernel=False,
                 n_fft=256, hop_length=16, device='cuda:0'):
        super().__init__()
        _cabi.lib()   # fail at construction time if the extension is missing
        self.wavelength = torch.nn.Parameter(torch.as_tensor(wavelength, dtype=torch.float32),
                                             requires_grad=bool(train_wavelength))
        self.radar_location = torch.nn.Parameter(torch.as_tensor(radar_location, dtype=torch.float32),
                                                 requires_grad=bool(train_radar_location))
        self.src, self.dst = map(list, zip(*edges))
        self.stft = _STFTKernels(n_fft, hop_length, bool(train_stft_kernel), device)
        self.n_fft = n_fft
        self.hop_length = hop_length
        self._src_c = _cabi.i32_array(self.src)
        self._dst_c = _cabi.i32_array(self.dst)
        # Opt-in for streams of INDEPENDENT batches (VR_FLAG_INPUTS_READY): set True only if the kernel launched just
        # before each forward on the same stream neither produces x nor touches the output; consecutive forwards then
        # overlap (a batch starts in the SM slots the previous one has vacated).  Default: plain stream order.
        self.assume_inputs_ready = False
        # Profiling aid: True wraps every launch of this layer in an NVTX range ("VirtualRadar.<entry>") so that timeline
        # tools show the fused stage among the model's other kernels (tracing hooks).  Off by default.
        self.nvtx_ranges = False

    # the ctypes arrays are not picklable / deep-copyable (DataParallel.replicate copies __dict__)
    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_src_c", None)
        d.pop("_dst_c", None)
        d.pop("_host_params", None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._src_c = _cabi.i32_array(self.src)
        self._dst_c = _cabi.i32_array(self.dst)

    def _check_stft(self):
        """The fused kernel evaluates the analytic Hann-windowed DFT with an FFT.  Trained `stft.wsin/wcos` (reference
        `train_stft_kernel=True`, or a checkpoint of such a model) must never silently be replaced by it, so the
        parameters are compared with the analytic kernels (one small device-to-host copy) whenever they may have
        changed: the verdict is cached against the tensors' identity and version counters, which every in-place
        update through the parameters -- optimizer step, `load_state_dict`, `wsin.copy_()`, `.to()` -- changes.  (Writes
        through `wsin.data` bypass PyTorch's version counters; call `invalidate_stft_cache()` after such an edit.)"""
        if getattr(self, "_stft_trusted", False):       # a DataParallel replica: the parent module has just checked
            return
        key = tuple((t.data_ptr(), t._version, t.device) for t in (self.stft.wsin, self.stft.wcos))
        if getattr(self, "_stft_key", None) != key:
            self._stft_is_dft = self.stft.is_dft()
            self._stft_key = key

    def invalidate_stft_cache(self):
        self._stft_key = None

    def _replicate_for_data_parallel(self):
        if not (self.stft.wsin.requires_grad or self.stft.wcos.requires_grad):
            self._check_stft()                          # once on the parent instead of once per replica and forward
        replica = super()._replicate_for_data_parallel()
        replica._stft_trusted = True
        return replica

    def _general_stft(self):
        """True -> CUDA synthesis, then the STFT against `wsin`/`wcos` as they are (`_STFTKernels.logmag`).  Trainable
        kernels always take this path, with or without grad mode: an eval pass of a model being trained must see the
        kernels the optimizer has produced."""
        if self.stft.wsin.requires_grad or self.stft.wcos.requires_grad or self.n_fft != self._FUSED_N_FFT:
            return True
        self._check_stft()
        return not self._stft_is_dft

    def output_shape(self, x_shape):
        return (x_shape[0], self.n_fft, x_shape[2] // self.hop_length + 1)

    def _check_input(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() != 5 or x.shape[1] != 3:
            raise ValueError("expected x of shape (batch, 3, timesteps, vertices, num_graphs), got %s"
                             % (tuple(x.shape) if isinstance(x, torch.Tensor) else type(x),))
        if x.dtype != torch.float32:
            raise ValueError("VirtualRadar computes in float32 like the reference; got %s" % x.dtype)

    def _check_device(self, x):
        if not x.is_cuda:
            raise RuntimeError("VirtualRadar (B200) has no CPU path: move x to a CUDA device, or call "
                               "forward_host(x) to stream a pinned host batch through the GPU")
        lam, loc = self.wavelength, self.radar_location
        if lam.device != x.device or loc.device != x.device:
            raise RuntimeError("module parameters are on %s but x is on %s; call .to(x.device)" % (lam.device, x.device))
        return lam, loc

    def _prepare(self, x):
        """Pick the range rounding mode from the caller's strides BEFORE normalising the layout
        (ATen's CPU norm rounds differently when the coordinate axis is innermost)."""
        flags = _cabi.VR_FLAG_RANGE_FMA if x.stride(1) == 1 and x.shape[1] > 1 else 0
        return x.contiguous(), flags

    _FUSED_N_FFT = 256      # the fused kernel's FFT size; other n_fft values go through the general-kernel path

    def _launch(self, xc, flags, want_iq=False):
        N, _, T, V, M = xc.shape
        n_fft = self._FUSED_N_FFT if want_iq else self.n_fft       # iq does not depend on the STFT parameters
        if self.assume_inputs_ready:
            flags |= _cabi.VR_FLAG_INPUTS_READY
        out = torch.empty((N, n_fft, T // self.hop_length + 1), dtype=torch.float32, device=xc.device)
        iq = torch.empty((N, T, 2), dtype=torch.float32, device=xc.device) if want_iq else None
        if N == 0:
            return out, iq
        L = _cabi.lib()
        if self.nvtx_ranges:
            torch.cuda.nvtx.range_push("VirtualRadar.forward%s N=%d T=%d" % ("+iq" if want_iq else "", N, T))
        try:
            with torch.cuda.device(xc.device):
                stream = ctypes.c_void_p(torch.cuda.current_stream(xc.device).cuda_stream)
                args = (xc.data_ptr(), N, T, V, M, self._src_c, self._dst_c, len(self.src), self.wavelength.data_ptr(),
                        self.radar_location.data_ptr(), n_fft, self.hop_length, flags, out.data_ptr())
                rc = L.vr_forward_debug_f32(*args, iq.data_ptr(), stream) if want_iq else L.vr_forward_f32(*args, stream)
        finally:
            if self.nvtx_ranges:
                torch.cuda.nvtx.range_pop()
        _cabi.check(rc)
        return out, iq

    def _needs_grad(self, x=None):
        return torch.is_grad_enabled() and (self.wavelength.requires_grad or self.radar_location.requires_grad
                                            or (x is not None and x.requires_grad))

    def forward(self, x):
        self._check_input(x)
        self._check_device(x)
        xc, flags = self._prepare(x)
        return self._run(x, xc, flags)

    def _run(self, x, xc, flags):
        """forward() after the layout has been normalised: xc standard-contiguous, flags carry the range rounding mode."""
        lam, loc = self.wavelength, self.radar_location
        if self._general_stft():     # trained / trainable STFT kernels: synthesis on our kernels, STFT as a GEMM
            if xc.shape[2] <= max(self.n_fft, self._FUSED_N_FFT) // 2:
                raise ValueError("T=%d must exceed n_fft/2=%d: reflect padding needs it"
                                 % (xc.shape[2], max(self.n_fft, self._FUSED_N_FFT) // 2))
            if xc.shape[0] == 0:
                return self._launch(xc, flags)[0]
            iq = _SynthFunction.apply(lam, loc, xc, self, flags) if self._needs_grad(x) else self._launch(xc, flags, want_iq=True)[1]
            return self.stft.logmag(iq)
        if self._needs_grad(x) and xc.shape[0] > 0:
            return _RadarFunction.apply(lam, loc, xc, self, flags)
        return self._launch(xc, flags)[0]

    def forward_notebook(self, data, num_pad_frames=1, sigma=3):
        """virtual_radar_example.ipynb cells 2-4 on the device: `data` is one body's (T, V, 3) array (or a batch
        (N, T, V, 3)), float64 or float32 CUDA tensor.  Equals
        `self(torch.Tensor(np.expand_dims(utils.pad_frames(data, num_pad_frames, sigma).transpose(2, 0, 1), [0, -1])))`
        of the reference (utils.py:82-89 + layers/virtual_radar.py:79-134): the up-sampling kernel writes float32 straight
        in the layer's layout (C ABI vr_pad_frames_joints, planar), and the launch is told to round the radar range the way
        the reference does for the notebook's coordinate-innermost tensor (VR_FLAG_RANGE_FMA) -- no layout copy of the
        k-times larger array in between."""
        from ..upsample import pad_frames_notebook
        if isinstance(data, torch.Tensor) and data.shape[-1] != 3:
            raise ValueError("expected (T, V, 3) or (N, T, V, 3) joint positions, got %s" % (tuple(data.shape),))
        x = pad_frames_notebook(data, num_pad_frames, sigma, planar=True)
        self._check_input(x)
        self._check_device(x)
        return self._run(x, x, _cabi.VR_FLAG_RANGE_FMA)

    def forward_image(self, x, image_size=256):
        """The layer fused with its consumer's input stage (reference models/resnet.py:24-26):
        equals `F.interpolate(self(x).unsqueeze(1), image_size)` (nearest) bit for bit, in one launch
        that writes (N, 1, image_size, image_size) directly and transforms only the frames the
        resize keeps."""
        self._check_input(x)
        lam, loc = self._check_device(x)
        image_size = int(image_size)
        if image_size < 1:
            raise ValueError("image_size must be positive, got %d" % image_size)
        if self._needs_grad(x) or self._general_stft():     # gradients wanted / general STFT kernels: spectrogram, then torch's resize
            return torch.nn.functional.interpolate(self.forward(x).unsqueeze(1), image_size)
        xc, flags = self._prepare(x)
        if self.assume_inputs_ready:
            flags |= _cabi.VR_FLAG_INPUTS_READY
        N, _, T, V, M = xc.shape
        out = torch.empty((N, 1, image_size, image_size), dtype=torch.float32, device=x.device)
        if N == 0:
            return out
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            rc = _cabi.lib().vr_forward_image_f32(xc.data_ptr(), N, T, V, M, self._src_c, self._dst_c, len(self.src),
                                                  lam.data_ptr(), loc.data_ptr(), self.n_fft, self.hop_length,
                                                  flags, image_size, out.data_ptr(), ctypes.c_void_p(stream))
        _cabi.check(rc)
        return out

    def forward_upsampled(self, x, num_pad_frames=250, sigma=3, image_size=None):
        """The data loader's temporal up-sampling fused in front of the layer: x is the RAW
        (N,3,T,V,M) batch; the result equals `self(pad_frames(x, num_pad_frames, sigma))` -- or
        `self.forward_image(pad_frames(...), image_size)` when `image_size` is given -- bit for bit,
        without the `num_pad_frames`-times larger batch ever existing in HBM (C ABI
        vr_forward_upsampled_f32; replaces reference utils.py:128-140 + layers/virtual_radar.py:79-134
        [+ models/resnet.py:24-26])."""
        self._check_input(x)
        lam, loc = self._check_device(x)
        xc = x.contiguous()
        N, _, T, V, M = xc.shape
        k = int(num_pad_frames)
        img = 0 if image_size is None else int(image_size)
        if image_size is not None and img < 1:
            raise ValueError("image_size must be positive, got %d" % img)
        if x.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("the temporal up-sampling has no backward pass: detach x, or differentiate "
                                      "forward() on an already up-sampled batch")
        if self._needs_grad() or self._general_stft():      # trainable parameters / general STFT kernels: materialise the up-sampled batch
            from ..upsample import pad_frames
            up = pad_frames(xc, k, sigma)
            return self.forward_image(up, img) if img else self.forward(up)
        shape = (N, 1, img, img) if img else (N, self.n_fft, (k * T) // self.hop_length + 1)
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
        if N == 0:
            return out
        L = _cabi.lib()
        nbytes = int(L.vr_upsampled_workspace_bytes(N, T, V, M))
        work = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            # the up-sampled tensor the reference's loader builds is standard-contiguous: range rounding mode "seq"
            rc = L.vr_forward_upsampled_f32(xc.data_ptr(), N, T, V, M, self._src_c, self._dst_c, len(self.src),
                                            lam.data_ptr(), loc.data_ptr(), self.n_fft, self.hop_length, 0,
                                            k, ctypes.c_float(float(sigma)), img, work.data_ptr(), nbytes,
                                            out.data_ptr(), ctypes.c_void_p(stream))
        _cabi.check(rc)
        return out

    def forward_debug(self, x):
        """forward plus the intermediate complex baseband signal (N,T,2); for stage-level parity tests."""
        self._check_input(x)
        self._check_device(x)
        xc, flags = self._prepare(x)
        return self._launch(xc, flags, want_iq=True)

    def forward_host(self, x, out=None, sub_batch=0, device=None):
        """End-to-end call on HOST tensors: x (N,3,T,V,M) float32 on the CPU (pinned for full copy
        speed) -> log-spectrograms on the CPU.  Sub-batches are pipelined H2D / kernel / D2H inside
        the C library (vr_forward_host_f32)."""
        self._check_input(x)
        if x.is_cuda:
            raise ValueError("forward_host takes a CPU tensor; use forward() for CUDA tensors")
        if self._general_stft():
            raise NotImplementedError("forward_host evaluates the analytic Hann-windowed DFT only; stft.wsin/wcos are "
                                      "trained or trainable: use forward() on a CUDA tensor")
        xc, flags = self._prepare(x)
        N, _, T, V, M = xc.shape
        if out is None:
            out = torch.empty(self.output_shape(xc.shape), dtype=torch.float32, pin_memory=True)
        elif (not isinstance(out, torch.Tensor) or out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous()
              or tuple(out.shape) != tuple(self.output_shape(xc.shape))):
            raise ValueError("out must be a contiguous float32 CPU tensor of shape %s" % (tuple(self.output_shape(xc.shape)),))
        if N == 0:
            return out
        dev = torch.device(device) if device is not None else self.wavelength.device
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        lam, loc = self._host_parameters()
        with torch.cuda.device(dev):
            rc = _cabi.lib().vr_forward_host_f32(xc.data_ptr(), N, T, V, M, self._src_c, self._dst_c, len(self.src),
                                                 ctypes.c_float(lam), loc,
                                                 self.n_fft, self.hop_length, flags, out.data_ptr(), int(sub_batch))
        _cabi.check(rc)
        return out

    def _host_parameters(self):
        """Host copies of the two radar parameters for the host-buffer entry point, refreshed only when the parameters
        change (identity / version counters) -- not two device-to-host synchronisations per call."""
        key = tuple((t.data_ptr(), t._version, t.device) for t in (self.wavelength, self.radar_location))
        cached = getattr(self, "_host_params", None)
        if cached is None or cached[0] != key:
            lam = float(self.wavelength.detach().cpu())
            loc = (ctypes.c_float * 3)(*[float(v) for v in self.radar_location.detach().cpu().tolist()])
            cached = self._host_params = (key, lam, loc)
        return cached[1], cached[2]

    def extra_repr(self):
        return "bones=%d, n_fft=%d, hop_length=%d" % (len(self.src), self.n_fft, self.hop_length)
