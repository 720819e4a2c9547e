// vr_backward.cuh -- gradients of the VirtualRadar layer with respect to its two radar parameters
// (`wavelength`, `radar_location`; reference layers/virtual_radar.py:40-41, 65-69: the constructor's
// train_wavelength / train_radar_location flags make them trainable nn.Parameters and PyTorch autograd
// differentiates forward(), :79-134).  Two kernels:
//
//   vr_stft_adjoint_kernel   grad_out (N, 256, F) and the saved complex baseband signal z (N, T, 2)
//                            -> dL/dz (N, T, 2): per frame, X = FFT(hann * zp); G = g * X / (|X| (|X| + 1e-6))
//                            (log-magnitude and fftshift, :126-133); dL/dzp = hann * conj(FFT(conj(G))) (the
//                            adjoint of the windowed DFT, :124-125); overlap-add through the reflect padding.
//   vr_synth_adjoint_kernel  dL/dz and x -> dL/dwavelength, dL/dradar_location (and, on request, dL/dx): per (sequence, time step, body,
//                            bone) the forward geometry is recomputed (range and phase with the forward's exact
//                            float32 rounding, the rest in float64) and the analytic derivatives of
//                            z = sum amp * exp(j theta) are accumulated in float64.
//
// theta = 4 pi d / lambda reaches 1e4..1e5 rad, so dtheta/dlambda = -theta/lambda is of order 1e7..1e8: the
// reference's float32 autograd result carries relative errors of 1e-3 and more; these kernels are checked
// against the float64 autograd of the oracle (tests/test_backward_gpu.py).
#pragma once
#include "vr_kernels.cuh"

namespace vr {

struct BwdParams {
    const float* x;              // (N,3,T,V,M)
    const float* iq;             // (N,T,2) saved by the forward (vr_forward_debug_f32)
    const float* gout;           // (N,256,F)
    float* gz;                   // (N,T,2) work buffer, zeroed before the adjoint STFT
    double* gparams;             // [dL/dlambda, dL/dLx, dL/dLy, dL/dLz], accumulated with atomics
    float* gx;                   // optional dL/dx (N,3,T,V,M), zeroed before the launch; nullptr = not wanted
    const float* lam_ptr;
    const float* loc_ptr;
    long long N, T;
    int V, M, E, F, hop, VM;
    int fma_range;
    float inv_E;
    uint16_t src[NG * MAX_EG], dst[NG * MAX_EG];
};

#ifdef __CUDACC__
// 256-point forward DFT (e^{-j...}) of one frame by one warp.  In: v[q] = sample lane + 32 q.  Out: v holds
// the bins bwd_bin(lane, j), j = 0..7.  Same radix 8 x 8 x 4 decomposition as the forward kernel's STFT.
__device__ __forceinline__ int bwd_bin(int lane, int j) {
    const int k1 = lane >> 2, b4 = lane & 3;
    return k1 + 8 * (b4 + 4 * (j >> 2)) + 64 * (j & 3);
}
__device__ __forceinline__ void bwd_fft256(c2 (&v)[8], float2* __restrict__ xch, const float4* __restrict__ tw1,
                                           const float4* __restrict__ tw2, int lane) {
    const int k1 = lane >> 2, b4 = lane & 3;
    dft8(v);
#pragma unroll
    for (int q = 1; q < 8; ++q) v[q] = pcmul(v[q], tw1[(q - 1) * 32 + lane]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) xch[q * XCH_STRIDE + lane] = v[q];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 8; ++a) v[a] = xch[k1 * XCH_STRIDE + 4 * a + b4];
    dft8(v);
#pragma unroll
    for (int c = 1; c < 8; ++c) v[c] = pcmul(v[c], tw2[(c - 1) * 4 + b4]);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) xch[k1 * XCH_STRIDE + 4 * c + b4] = v[c];
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int c = b4 + 4 * hh;
        const float4* src4 = reinterpret_cast<const float4*>(&xch[k1 * XCH_STRIDE + 4 * c]);
        const float4 p01 = src4[0], p23 = src4[1];
        dft4(make_float2(p01.x, p01.y), make_float2(p01.z, p01.w), make_float2(p23.x, p23.y),
             make_float2(p23.z, p23.w), v[4 * hh], v[4 * hh + 1], v[4 * hh + 2], v[4 * hh + 3]);
    }
    __syncwarp();
}
__device__ __forceinline__ int bwd_reflect(int t, int T) {       // nnAudio center=True, pad_mode='reflect'
    t = t < 0 ? -t : t;
    return t >= T ? 2 * (T - 1) - t : t;
}

constexpr int BWD_WARPS = 8;

__global__ void __launch_bounds__(BWD_WARPS * 32) vr_stft_adjoint_kernel(const __grid_constant__ BwdParams p) {
    __shared__ float4 tw1[7 * 32 + 7 * 4];
    __shared__ float hann[NFFT];
    __shared__ float2 xch_all[BWD_WARPS][8 * XCH_STRIDE];
    __shared__ float2 tr_all[BWD_WARPS][NFFT];
    float4* tw2 = tw1 + 7 * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 7 * 32 + 7 * 4; i += blockDim.x) {
        const int e = i < 7 * 32 ? (i & 31) * ((i >> 5) + 1) : 8 * ((i - 7 * 32) & 3) * (((i - 7 * 32) >> 2) + 1);
        float sn, cs;
        sincospif((float)(e & 255) * (2.0f / NFFT), &sn, &cs);
        tw1[i] = make_float4(cs, -sn, sn, cs);
    }
    for (int i = tid; i < NFFT; i += blockDim.x) hann[i] = fmaf(-0.5f, cospif((float)i * (2.0f / NFFT)), 0.5f);
    __syncthreads();
    float2* xch = xch_all[warp];
    float2* tr = tr_all[warp];
    const int T = (int)p.T;
    const long long frames = p.N * (long long)p.F;
    for (long long fr = (long long)blockIdx.x * BWD_WARPS + warp; fr < frames; fr += (long long)gridDim.x * BWD_WARPS) {
        const long long n = fr / p.F;
        const int f = (int)(fr - n * p.F);
        const float2* z = reinterpret_cast<const float2*>(p.iq) + n * T;
        const int fstart = f * p.hop - NFFT / 2;
        c2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int nidx = lane + 32 * q;
            v[q] = pscale(__ldg(z + bwd_reflect(fstart + nidx, T)), hann[nidx]);
        }
        bwd_fft256(v, xch, tw1, tw2, lane);
        // log-magnitude + fftshift backward: G = g X / (|X| (|X| + 1e-6)); park conj(G) by bin
        const float* g = p.gout + n * (long long)NFFT * p.F + f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kbin = bwd_bin(lane, j);
            const int row = (kbin + NFFT / 2) & (NFFT - 1);
            const float a = sqrtf(v[j].x * v[j].x + v[j].y * v[j].y);
            const float s = a > 0.f ? __ldg(g + (size_t)row * p.F) / (a * (a + 1e-6f)) : 0.f;
            tr[kbin] = make_float2(s * v[j].x, -s * v[j].y);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = tr[lane + 32 * q];
        bwd_fft256(v, xch, tw1, tw2, lane);
        float* gz = p.gz + n * (long long)T * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int nidx = bwd_bin(lane, j);
            const int t = bwd_reflect(fstart + nidx, T);
            const float w = hann[nidx];
            atomicAdd(gz + 2 * t, w * v[j].x);
            atomicAdd(gz + 2 * t + 1, -w * v[j].y);
        }
    }
}

// exact float32 range and phase of a joint, as the forward computes them
__device__ __forceinline__ void bwd_range_phase(float jx, float jy, float jz, bool fma_range, float lam, float lam_rcp,
                                                float& d, float& th) {
    const float d2 = fma_range ? __fmaf_rn(jz, jz, __fmaf_rn(jy, jy, __fmul_rn(jx, jx)))
                               : __fadd_rn(__fadd_rn(__fmul_rn(jx, jx), __fmul_rn(jy, jy)), __fmul_rn(jz, jz));
    d = sqrt_rn_fast(d2);
    th = div_rn_fast(__fmul_rn(12.566370614359172f, d), lam, lam_rcp);
}

__global__ void __launch_bounds__(128) vr_synth_adjoint_kernel(const __grid_constant__ BwdParams p) {
    const int T = (int)p.T;
    const long long total = p.N * (long long)T;
    const float lam = __ldg(p.lam_ptr);
    const float Lx = __ldg(p.loc_ptr), Ly = __ldg(p.loc_ptr + 1), Lz = __ldg(p.loc_ptr + 2);
    const float lam_rcp = rcp_refined(lam);
    const double PI = 3.14159265358979323846;
    double g_lam = 0.0, g_lx = 0.0, g_ly = 0.0, g_lz = 0.0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / T;
        const int t = (int)(idx - n * T);
        const double gI = (double)__ldg(p.gz + idx * 2), gQ = (double)__ldg(p.gz + idx * 2 + 1);
        if (gI == 0.0 && gQ == 0.0) continue;
        const float* x0 = p.x + (n * 3 * T + t) * (long long)p.VM;       // x-plane row of this time step
        const long long ps = (long long)T * p.VM;                         // floats between coordinate planes
        for (int m = 0; m < p.M; ++m) {
            // mean bone length (:110-113), float32 like the forward
            float sumB = 0.f;
            for (int e = 0; e < p.E; ++e) {
                const float* s = x0 + p.src[e] * p.M + m;
                const float* d = x0 + p.dst[e] * p.M + m;
                const float bx = __fsub_rn(__ldg(d), __ldg(s)), by = __fsub_rn(__ldg(d + ps), __ldg(s + ps)),
                            bz = __fsub_rn(__ldg(d + 2 * ps), __ldg(s + 2 * ps));
                const float bb = p.fma_range ? __fmaf_rn(bz, bz, __fmaf_rn(by, by, __fmul_rn(bx, bx)))
                                             : __fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz));
                sumB = __fadd_rn(sumB, sqrt_rn_fast(bb));
            }
            if (sumB == 0.f) continue;                                    // absent body: contributes exactly 0
            const double cbar = (double)__fmul_rn(sumB, p.inv_E);
            const double c = cbar * cbar, K = sqrt(PI) * cbar;
            float* gx0 = p.gx ? p.gx + (n * 3 * T + t) * (long long)p.VM : nullptr;   // this thread's own rows of dL/dx
            double G_c = 0.0;                                              // dL/d(mean bone length), over this body's bones
            for (int e = 0; e < p.E; ++e) {
                const float* s = x0 + p.src[e] * p.M + m;
                const float* d = x0 + p.dst[e] * p.M + m;
                const float sx = __ldg(s), sy = __ldg(s + ps), sz = __ldg(s + 2 * ps);
                const float dx = __ldg(d), dy = __ldg(d + ps), dz = __ldg(d + 2 * ps);
                // range, phase: the forward's float32 values
                float rng, th;
                bwd_range_phase(__fsub_rn(sx, Lx), __fsub_rn(sy, Ly), __fsub_rn(sz, Lz), p.fma_range != 0, lam, lam_rcp, rng, th);
                double sn, cs;
                sincos((double)th, &sn, &cs);
                // aspect cosine and amplitude, float64 from the float32 inputs
                const double Bx = (double)dx - sx, By = (double)dy - sy, Bz = (double)dz - sz;
                const double Ax = (double)Lx - 0.5 * ((double)sx + dx), Ay = (double)Ly - 0.5 * ((double)sy + dy),
                             Az = (double)Lz - 0.5 * ((double)sz + dz);
                const double na = sqrt(Ax * Ax + Ay * Ay + Az * Az), nb = sqrt(Bx * Bx + By * By + Bz * Bz);
                const double dot = Ax * Bx + Ay * By + Az * Bz;
                const double q = na * nb + 1e-6;
                const double u = dot / q;
                const double den = 1.0 + (c - 1.0) * u * u;
                const double amp = K / den;
                const double dL_dth = amp * (gQ * cs - gI * sn);
                const double dL_damp = gI * cs + gQ * sn;
                g_lam += dL_dth * (-(double)th / (double)lam);
                double gsx = 0.0, gsy = 0.0, gsz = 0.0, gdx = 0.0, gdy = 0.0, gdz = 0.0;   // dL/dS, dL/dD of this bone
                if (rng > 0.f) {
                    const double kth = dL_dth * (4.0 * PI / (double)lam) / (double)rng;
                    g_lx += kth * ((double)Lx - sx); g_ly += kth * ((double)Ly - sy); g_lz += kth * ((double)Lz - sz);
                    gsx = kth * ((double)sx - Lx); gsy = kth * ((double)sy - Ly); gsz = kth * ((double)sz - Lz);
                }
                const double kamp = dL_damp * (-K * 2.0 * u * (c - 1.0) / (den * den));
                if (na > 0.0) {
                    const double r2 = dot * nb / (na * q * q);
                    const double ax = kamp * (Bx / q - r2 * Ax), ay = kamp * (By / q - r2 * Ay), az = kamp * (Bz / q - r2 * Az);   // dL/dA
                    g_lx += ax; g_ly += ay; g_lz += az;
                    gsx -= 0.5 * ax; gsy -= 0.5 * ay; gsz -= 0.5 * az;
                    gdx -= 0.5 * ax; gdy -= 0.5 * ay; gdz -= 0.5 * az;
                }
                if (gx0) {
                    if (nb > 0.0) {
                        const double r3 = dot * na / (nb * q * q);
                        const double bx = kamp * (Ax / q - r3 * Bx), by = kamp * (Ay / q - r3 * By), bz = kamp * (Az / q - r3 * Bz);  // dL/dB via u
                        gsx -= bx; gsy -= by; gsz -= bz;
                        gdx += bx; gdy += by; gdz += bz;
                    }
                    G_c += dL_damp * sqrt(PI) * (1.0 / den - 2.0 * c * u * u / (den * den));
                    float* gs = gx0 + p.src[e] * p.M + m;
                    float* gd = gx0 + p.dst[e] * p.M + m;
                    gs[0] += (float)gsx; gs[ps] += (float)gsy; gs[2 * ps] += (float)gsz;
                    gd[0] += (float)gdx; gd[ps] += (float)gdy; gd[2 * ps] += (float)gdz;
                }
            }
            if (gx0) {                                                     // through the mean bone length (:110-113)
                for (int e = 0; e < p.E; ++e) {
                    const float* s = x0 + p.src[e] * p.M + m;
                    const float* d = x0 + p.dst[e] * p.M + m;
                    const double Bx = (double)__ldg(d) - __ldg(s), By = (double)__ldg(d + ps) - __ldg(s + ps),
                                 Bz = (double)__ldg(d + 2 * ps) - __ldg(s + 2 * ps);
                    const double nb = sqrt(Bx * Bx + By * By + Bz * Bz);
                    if (nb > 0.0) {
                        const double kc = G_c * (double)p.inv_E / nb;
                        float* gs = gx0 + p.src[e] * p.M + m;
                        float* gd = gx0 + p.dst[e] * p.M + m;
                        gs[0] -= (float)(kc * Bx); gs[ps] -= (float)(kc * By); gs[2 * ps] -= (float)(kc * Bz);
                        gd[0] += (float)(kc * Bx); gd[ps] += (float)(kc * By); gd[2 * ps] += (float)(kc * Bz);
                    }
                }
            }
        }
    }
    // block reduction, then one float64 atomic per block and parameter
    __shared__ double red[4][4];
    double vals[4] = {g_lam, g_lx, g_ly, g_lz};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double v = vals[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const double v = (red[threadIdx.x][0] + red[threadIdx.x][1]) + (red[threadIdx.x][2] + red[threadIdx.x][3]);
        if (v != 0.0) atomicAdd(p.gparams + threadIdx.x, v);
    }
}
#endif  // __CUDACC__

}  // namespace vr
