/* virtual_radar_b200.h -- C ABI of the B200-native VirtualRadar hot path.
 *
 * The reference (itskalvik/skeleton-action-recognition) is pure Python/PyTorch and has no FFI of
 * its own; this header is the boundary a maintainer would bind (ctypes stub in INTEGRATION.md) to
 * replace the body of `VirtualRadar.forward` (reference layers/virtual_radar.py:79-134) and the
 * nnAudio `STFT` it constructs (layers/virtual_radar.py:71-76, called at :124-125).
 *
 * Conventions
 *   - plain C, no torch types; all `*_dev` pointers are device pointers on the CURRENT CUDA device,
 *     all `*_host` pointers are host pointers.  The caller owns every buffer.
 *   - x      : (N, 3, T, V, M) float32, standard-contiguous  -- the layer's input layout
 *              (layers/virtual_radar.py:82-83 docstring; produced by data_gen/gen_joint_data.py).
 *   - out    : (N, n_fft, T / hop + 1) float32, contiguous   -- layers/virtual_radar.py:131-134.
 *   - edges  : E bones as (src[e], dst[e]) joint indices, HOST arrays (the reference keeps them as
 *              Python lists `self.src`, `self.dst`, layers/virtual_radar.py:70).
 *   - wavelength_dev (1 float), radar_loc_dev (3 floats): the layer's two nn.Parameters
 *              (layers/virtual_radar.py:65-69) read on the device, so no host sync is needed.
 *   - every entry point returns 0 or a negative VR_ERR_* code; vr_last_error() gives the message
 *     for the calling thread.  Nothing throws, exits, or synchronises the device, and work is
 *     only enqueued on `stream` (a cudaStream_t passed as void*).
 *   - re-entrant: no global mutable state except a per-thread error string and a per-device
 *     one-time kernel attribute setup.
 */
#ifndef VIRTUAL_RADAR_B200_H
#define VIRTUAL_RADAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_ABI_VERSION 1

#define VR_OK               0
#define VR_ERR_ARG         -1   /* null pointer, bad flag                       (ValueError)   */
#define VR_ERR_SHAPE       -2   /* T <= n_fft/2, edge index >= V, N/T/V/M <= 0   (ValueError)   */
#define VR_ERR_UNSUPPORTED -3   /* shape outside what the kernels cover          (NotImplementedError) */
#define VR_ERR_CUDA        -4   /* CUDA runtime error                            (RuntimeError) */

/* flags */
#define VR_FLAG_RANGE_FMA  1u   /* round the radar->joint range like ATen does when the caller's
                                   tensor had the coordinate axis innermost (x.stride(1)==1):
                                   sqrt(fma(c,c,fma(b,b,a*a))) instead of sqrt((a*a+b*b)+c*c).
                                   Replaces torch.norm(dim=1) at layers/virtual_radar.py:99.      */

#define VR_FLAG_INPUTS_READY 2u  /* Caller's guarantee for streams of INDEPENDENT batches: the kernel launched just
                                   before this one on `stream` does not write anything this call reads (x, the two
                                   parameters) and does not touch `out`.  The launches are programmatic dependent
                                   launches; with the flag this batch's reads no longer wait for the previous kernel
                                   to finish, so it starts in the SM slots that kernel has already vacated (its writes
                                   still wait).  Without the flag (the default, and what the torch module passes)
                                   plain stream order holds.                                                       */

int         vr_abi_version(void);
const char* vr_last_error(void);

/* The whole layer in one fused launch: geometry + RCS + baseband synthesis + windowed 256-point
 * STFT + log magnitude + fftshift.  Replaces VirtualRadar.forward, layers/virtual_radar.py:79-134.
 * n_fft must be 256 in this ABI version (the reference default, layers/virtual_radar.py:43).      */
int vr_forward_f32(const float* x_dev, int64_t N, int64_t T, int32_t V, int32_t M,
                   const int32_t* src_host, const int32_t* dst_host, int32_t E,
                   const float* wavelength_dev, const float* radar_loc_dev,
                   int32_t n_fft, int32_t hop, uint32_t flags,
                   float* out_dev, void* stream);

/* Same launch, additionally writing the intermediate complex baseband signal
 * (N, T, 2) = [I, Q] -- `phase_data` after the sum at layers/virtual_radar.py:123 -- for stage-level
 * parity tests.  The product path never materialises it.                                           */
int vr_forward_debug_f32(const float* x_dev, int64_t N, int64_t T, int32_t V, int32_t M,
                         const int32_t* src_host, const int32_t* dst_host, int32_t E,
                         const float* wavelength_dev, const float* radar_loc_dev,
                         int32_t n_fft, int32_t hop, uint32_t flags,
                         float* out_dev, float* iq_dev, void* stream);

/* The layer fused with its consumer's input stage: reference models/resnet.py:24-26 applies
 * `x.unsqueeze(1)` and `torch.nn.functional.interpolate(x, image_size)` (mode 'nearest') to the
 * layer's output before the ResNet.  This entry writes that (N, 1, image_size, image_size) float32
 * image directly: image row r shows spectrogram row min(floor(r * float(n_fft)/image_size), n_fft-1),
 * image column c shows STFT frame min(floor(c * float(F)/image_size), F-1), F = T/hop + 1 (ATen's
 * legacy nearest indexing).  Only the frames the resize keeps are transformed (for T = 75 000 that is
 * 256 of 4 688), and the (N, n_fft, F) spectrogram never goes to HBM.  Values are bit-identical to
 * vr_forward_f32 followed by the resize.                                                          */
int vr_forward_image_f32(const float* x_dev, int64_t N, int64_t T, int32_t V, int32_t M,
                         const int32_t* src_host, const int32_t* dst_host, int32_t E,
                         const float* wavelength_dev, const float* radar_loc_dev,
                         int32_t n_fft, int32_t hop, uint32_t flags, int32_t image_size,
                         float* out_dev, void* stream);

/* End-to-end entry with HOST buffers (x_host pinned for full speed): splits the batch into
 * sub-batches and pipelines H2D copy / fused kernel / D2H copy on two internal streams of the
 * current device.  Blocks until out_host is complete.  Staging buffers are owned by the library,
 * cached per device, and released by vr_release_host_staging().                                    */
int vr_forward_host_f32(const float* x_host, int64_t N, int64_t T, int32_t V, int32_t M,
                        const int32_t* src_host, const int32_t* dst_host, int32_t E,
                        float wavelength, const float* radar_loc_host,
                        int32_t n_fft, int32_t hop, uint32_t flags,
                        float* out_host, int64_t sub_batch);
int vr_release_host_staging(void);

/* Temporal up-sampling of the joint trajectories to the radar sampling rate, the data loader's pre-stage:
 * replaces `Dataset.pad_frames` (reference utils.py:134-140: scipy gaussian_filter1d(sigma) along time,
 * then not-a-knot cubic interp1d to num_pad_frames*T frames, float64) together with the FloatTensor cast of
 * `Dataset.__getitem__` (utils.py:128-132), for a whole batch on the device.
 *   x_dev (N,3,T,V,M) float32 -> out_dev (N,3,num_pad_frames*T,V,M) float32, both contiguous; T >= 4.
 * The reference default is num_pad_frames=250, sigma=3 (utils.py:105).                              */
int vr_pad_frames_f32(const float* x_dev, int64_t N, int64_t T, int32_t V, int32_t M, int32_t num_pad_frames,
                      float sigma, float* out_dev, void* stream);

/* The notebook's variant of the up-sampling: replaces `utils.pad_frames` (reference utils.py:82-89, used by
 * virtual_radar_example.ipynb cells 2-4 on one body's (T, V, C) array): scipy gaussian_filter1d(sigma) along the JOINT
 * axis (axis=1 -- a quirk of the reference that is kept), then the not-a-knot cubic interp1d in time to
 * num_pad_frames*T frames in float64, then the float32 cast of torch.Tensor(...).
 *   x_dev (N, T, V, C) float64 (x_is_f64 != 0) or float32, contiguous; T >= 4, T <= 8500 (shared-memory spline solve).
 *   out_dev float32: planar_out == 0: (N, num_pad_frames*T, V, C) -- permuted to (N, C, k*T, V, 1) this is the notebook's
 *   coordinate-innermost tensor (strides (.., 1, V*C, C, C)), for which the layer picks VR_FLAG_RANGE_FMA itself;
 *   planar_out != 0: (N, C, num_pad_frames*T, V) -- the layer's own input layout with M = 1, for a direct vr_forward_f32
 *   call with VR_FLAG_RANGE_FMA and no layout copy in between.                                                        */
int vr_pad_frames_joints(const void* x_dev, int32_t x_is_f64, int64_t N, int64_t T, int32_t V, int32_t C,
                         int32_t num_pad_frames, float sigma, int32_t planar_out, float* out_dev, void* stream);

/* The data loader's up-sampling fused in front of the layer: equals vr_pad_frames_f32 followed by
 * vr_forward_f32 (image_size == 0; out_dev is (N, n_fft, num_pad_frames*T/hop + 1)) or by
 * vr_forward_image_f32 (image_size > 0; out_dev is (N, 1, image_size, image_size)) bit for bit, but the
 * (N,3,num_pad_frames*T,V,M) batch -- 45 MB per NTU sequence at the reference's 250 -- never exists:
 * a small launch smooths the raw trajectories and solves the spline (the per-interval cubics go to
 * `workspace_dev`, vr_upsampled_workspace_bytes(N,T,V,M) bytes, 1.4 MB per NTU sequence), and the radar
 * kernel evaluates every 32-step chunk from them in shared memory, in float64, rounding to float32
 * exactly where `Dataset.__getitem__` does (reference utils.py:128-140).  x_dev is the RAW (N,3,T,V,M)
 * batch.  The up-sampled tensor the reference builds is standard-contiguous, so pass flags = 0.     */
int64_t vr_upsampled_workspace_bytes(int64_t N, int64_t T, int32_t V, int32_t M);
int vr_forward_upsampled_f32(const float* x_dev, int64_t N, int64_t T, int32_t V, int32_t M,
                             const int32_t* src_host, const int32_t* dst_host, int32_t E,
                             const float* wavelength_dev, const float* radar_loc_dev,
                             int32_t n_fft, int32_t hop, uint32_t flags,
                             int32_t num_pad_frames, float sigma, int32_t image_size,
                             void* workspace_dev, int64_t workspace_bytes, float* out_dev, void* stream);

/* Backward pass: gradients of the layer's two radar parameters (the constructor's train_wavelength /
 * train_radar_location flags, reference layers/virtual_radar.py:40-41, 65-69; in the reference PyTorch
 * autograd differentiates forward()).  Inputs: the forward's x, the complex baseband signal the forward
 * saved (iq_dev, (N,T,2), from vr_forward_debug_f32) and grad_out_dev (N, n_fft, T/hop+1) = dL/d(out).
 * gz_work_dev: (N,T,2) float32 scratch (zeroed here; holds dL/d(iq) on return).  grad_params_dev: 4
 * float64 values ACCUMULATED (+=) with [dL/dwavelength, dL/dradar_location[0..2]]; the caller zeroes
 * them.  grad_x_dev (optional): gradient with respect to the skeleton data.  Two launches on `stream`: adjoint STFT (FFT, log-magnitude and fftshift backward, inverse
 * FFT, overlap-add through the reflect padding) and adjoint synthesis (float64 accumulation).        */
int vr_backward_f32(const float* x_dev, const float* iq_dev, const float* grad_out_dev,
                    int64_t N, int64_t T, int32_t V, int32_t M,
                    const int32_t* src_host, const int32_t* dst_host, int32_t E,
                    const float* wavelength_dev, const float* radar_loc_dev,
                    int32_t n_fft, int32_t hop, uint32_t flags,
                    float* gz_work_dev, double* grad_params_dev,
                    float* grad_x_dev /* (N,3,T,V,M) dL/dx, overwritten; NULL = not wanted */, void* stream);
/* Only the second stage: the caller supplies dL/d(iq) (grad_iq_dev, (N,T,2)) -- used when the STFT is a general,
 * trainable (n_fft x n_fft) kernel pair evaluated and differentiated outside this library (train_stft_kernel=True,
 * reference layers/virtual_radar.py:42,75).                                                                       */
int vr_synth_adjoint_f32(const float* x_dev, const float* grad_iq_dev, int64_t N, int64_t T, int32_t V, int32_t M,
                         const int32_t* src_host, const int32_t* dst_host, int32_t E,
                         const float* wavelength_dev, const float* radar_loc_dev, uint32_t flags,
                         double* grad_params_dev, float* grad_x_dev, void* stream);
/* the same without dL/dx */
int vr_backward_params_f32(const float* x_dev, const float* iq_dev, const float* grad_out_dev,
                           int64_t N, int64_t T, int32_t V, int32_t M,
                           const int32_t* src_host, const int32_t* dst_host, int32_t E,
                           const float* wavelength_dev, const float* radar_loc_dev,
                           int32_t n_fft, int32_t hop, uint32_t flags,
                           float* gz_work_dev, double* grad_params_dev, void* stream);

/* The STFT against GENERAL kernels (the constructor's train_stft_kernel=True, reference layers/virtual_radar.py:42,75,
 * or a checkpoint whose stft.wsin / stft.wcos are no longer the Hann-windowed Fourier kernels): what the reference
 * computes at layers/virtual_radar.py:124-133 with nnAudio's conv1d STFT, as ONE float32-accurate GEMM per batch on the
 * tensor cores (tcgen05.mma kind::tf32 with the error-compensated 3-term split, accumulator in tensor memory) whose
 * epilogue takes the magnitude, the log and the fftshift roll.  Any power-of-two n_fft in [16, 1024].
 *   iq_dev (N, T, 2): the complex baseband signal (vr_forward_debug_f32's iq_dev);  wsin_dev / wcos_dev (n_fft, 1, n_fft);
 *   out_dev (N, n_fft, T/hop + 1).  Work buffers (sizes from vr_stft_general_workspace_floats: parts[0] the
 *   reflect-padded planar signal -- the frame matrix is never materialised, the GEMM reads the frames as views of it --
 *   parts[1] kernel matrix, parts[2] saved Re/Im): frames_work and bt_work are scratch that the backward pass re-reads;
 *   c_save (optional, NULL = inference) keeps Re / Im of every bin for it.  bt_work, c_save (and dc_work below) must be
 *   16-byte aligned, iq_dev (and grad_iq_dev) 8-byte aligned -- VR_ERR_ARG otherwise: bulk copies, 16-byte accesses.  */
int64_t vr_stft_general_workspace_floats(int64_t N, int64_t T, int32_t n_fft, int32_t hop, int64_t parts[3]);
int vr_stft_general_f32(const float* iq_dev, int64_t N, int64_t T, int32_t n_fft, int32_t hop,
                        const float* wsin_dev, const float* wcos_dev, float* frames_work, float* bt_work,
                        float* c_save, float* out_dev, void* stream);
/* Its backward pass (the reference differentiates the conv1d graph with autograd): from dL/d(out) and the three buffers of
 * the forward to dL/d(iq) (N, T, 2) -- feed it to vr_synth_adjoint_f32 -- and dL/d(wsin), dL/d(wcos) (n_fft, 1, n_fft).
 * Two more GEMMs on the same kernel (dA = dC . Bt, dBt = dC^T . A).  dc_work: parts[2] floats; da_work: the frame
 * gradients, N * (T/hop + 1) * 2 * n_fft floats
 * (NULL together with grad_iq_dev when x and the radar parameters need no gradient); dbt_work: parts[1] floats (NULL
 * together with grad_wsin_dev / grad_wcos_dev when the kernels are frozen).                                         */
int vr_stft_general_backward_f32(const float* grad_out_dev, const float* frames_work, const float* bt_work, const float* c_save,
                                 int64_t N, int64_t T, int32_t n_fft, int32_t hop,
                                 float* dc_work, float* da_work, float* dbt_work,
                                 float* grad_iq_dev, float* grad_wsin_dev, float* grad_wcos_dev, void* stream);

/* Host-side planning, callable without a GPU (used by tests and by bench.py's reporting).
 * vr_plan fills `plan[16]`:
 *   [0] grid  [1] block  [2] dynamic smem bytes  [3] ring stages  [4] frames per job
 *   [5] jobs per sequence  [6] frames per output sub-batch  [7] uses TMA loads (0/1)
 *   [8] uses TMA bulk store (0/1)  [9] chunks per job  [10] max bones per lane group
 *   [11] max source joints per lane group  [12] z buffer capacity (samples)
 *   [13] CTAs per SM targeted [14] time steps per chunk  [15] lane groups                        */
int vr_plan(int64_t N, int64_t T, int32_t V, int32_t M,
            const int32_t* src_host, const int32_t* dst_host, int32_t E,
            int32_t n_fft, int32_t hop, int32_t sm_count, int64_t plan[16]);

/* vr_plan for vr_forward_image_f32; same layout except [4] output columns per job, [10] 1 if the
 * resize keeps fewer frames than the STFT has (only kept frames are transformed), [11] output columns. */
int vr_plan_image(int64_t N, int64_t T, int32_t V, int32_t M,
                  const int32_t* src_host, const int32_t* dst_host, int32_t E,
                  int32_t n_fft, int32_t hop, int32_t image_size, int32_t sm_count, int64_t plan[16]);

/* Host-side plan of the team-job schedule (see vr_set_schedule) for vr_forward_f32, callable without a GPU:
 *   [0] grid  [1] block (8 synthesis warps + one producer warp per team)  [2] dynamic smem bytes  [3] ring stages per
 *   team  [4] teams per CTA  [5] bytes per stage (max of a chunk's three planes and the output tile, which is built in
 *   the stage of a job's last chunk)  [6] bytes between the teams' z planes  [7] 1 if the automatic schedule picks it
 *   for this batch size.  VR_ERR_UNSUPPORTED if the shape does not qualify.                                          */
int vr_plan_team(int64_t N, int64_t T, int32_t V, int32_t M,
                 const int32_t* src_host, const int32_t* dst_host, int32_t E,
                 int32_t n_fft, int32_t hop, int32_t sm_count, int64_t plan[8]);

/* Host-side view of one job of a launch (test infrastructure for the job split shared by host and device):
 * geom = [sequence, first output column, columns, first frame, frames spanned, first source sample (chunk aligned),
 * last source sample, chunks].  image_size = 0 for vr_forward_f32's plan.                                          */
int vr_job_geometry(int64_t N, int64_t T, int32_t V, int32_t M,
                    const int32_t* src_host, const int32_t* dst_host, int32_t E,
                    int32_t n_fft, int32_t hop, int32_t image_size, int64_t job, int64_t geom[8]);

/* Bone -> lane-group assignment used by the synthesis stage (4 groups; all bones sharing a source
 * joint stay in one group so the range phase of that joint is evaluated once).  group_of_edge[E]. */
int vr_partition_edges(const int32_t* src_host, const int32_t* dst_host, int32_t E, int32_t V,
                       int32_t* group_of_edge);

/* GPU self-test (synchronous; test infrastructure): the kernels replace __fsqrt_rn/__fdiv_rn by their
 * branch-free fast-path sequences; this counts bitwise mismatches against the intrinsics over n
 * pseudo-random operands: [0] sqrt, [1] divide by `wavelength`, [2] general divide.               */
int vr_selftest_rounding(uint64_t n, float wavelength, uint64_t mismatches[3]);

/* Benchmark/tuning knob (process-wide, 0 = library default): warps per CTA (<=12), CTAs per SM
 * targeted (<=4), cap on TMA ring stages.  Not needed for normal use.                             */
int vr_set_tuning(int warps, int ctas_per_sm, int stages);

/* Benchmark knob (process-wide): which of the two schedules of the fused forward kernel a launch takes.  -1 (default)
 * automatic: batches with at least three sequences per team slot (592 on a B200), and all launches flagged
 * VR_FLAG_INPUTS_READY, run the team-job schedule -- every team of 4 warps owns a whole sequence, no CTA-wide
 * barriers -- smaller stream-ordered ones the cooperative schedule (both teams of a CTA share a sequence: half the
 * latency per sequence); 0: always cooperative; 1: team jobs whenever the shape qualifies (one job per sequence,
 * plain output, TMA-loadable chunks).  The results are bit-identical.                                              */
int vr_set_schedule(int mode);

/* Profiling aid (process-wide; NULL = off, the default): when set, every CTA of subsequent launches
 * writes 8 uint64 to dev_buf[8*blockIdx.x ..]: %globaltimer (ns) at [0] entry, [1] prologue done,
 * [2] first chunk landed, [3] first job's synthesis done, [4] first job's STFT done, [5] exit;
 * [7] = %smid.  The buffer must hold 8 * grid entries (grid from vr_plan).                        */
int vr_set_timeline_buffer(void* dev_u64_8_per_cta);

#ifdef __cplusplus
}
#endif
#endif /* VIRTUAL_RADAR_B200_H */
