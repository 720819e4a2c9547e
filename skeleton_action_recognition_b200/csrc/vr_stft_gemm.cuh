// vr_stft_gemm.cuh -- the STFT against GENERAL (trained / trainable) kernels, on the 5th-generation tensor cores.
//
// Reference: `train_stft_kernel=True` (layers/virtual_radar.py:42,75) makes nnAudio's Fourier kernels `stft.wsin`,
// `stft.wcos` (n_fft x 1 x n_fft each) parameters; the forward is then two conv1d per signal against whatever those
// tensors hold (layers/virtual_radar.py:124-129; nnAudio STFT.forward), so no FFT applies.  Written as
// one real GEMM per batch:
//
//     C[m, :] = A[m, :] . Bt^T,   m = (sequence, frame),  A[m] = [ I frame (n_fft) | Q frame (n_fft) ]   (K = 2 n_fft)
//     Bt rows: re_bin = [ wcos[bin] |  wsin[bin] ],  im_bin = [ -wsin[bin] | wcos[bin] ]                   (N = 2 n_fft)
//
// so that Re X = conv(I, wcos) + conv(Q, wsin), Im X = conv(Q, wcos) - conv(I, wsin) -- the combination of
// layers/virtual_radar.py:126-129 with nnAudio's imag = -conv(., wsin) -- come out of the same accumulator row.
//
// Kernel (`vr_gemm_tf32x3_kernel`): tcgen05.mma kind::tf32, M = 128 x N = 128 x K = 8 per instruction, accumulators in
// tensor memory, issued by one thread; operands staged in shared memory by the CTA's 256 threads in the canonical
// K-major no-swizzle layout (8 x 16-byte core matrices), two stages, released by tcgen05.commit on mbarriers.
// float32 accuracy from 10-bit TF32 mantissas by the error-compensated split a = hi + lo (hi = a rounded to TF32 with
// cvt.rna, lo = a - hi, exact): A.B = Ahi.Bhi + Alo.Bhi + Ahi.Blo -- three MMAs per K step; the dropped lo.lo term is
// 2^-22 relative.  The tensor core ACCUMULATES with truncation, one truncation per instruction: with all 3 x K/8 = 192
// instructions adding into one accumulator the first version was biased by ~4e-6 relative (measured, profiles/r02h).
// Hence FOUR accumulators in the 512 tensor-memory columns: the hi.hi products rotate over three of them (21
// truncations each) and the two small cross terms, 2^-11 of the result, go to the fourth, where truncation is
// harmless; the epilogue adds the four in float32.  Epilogue from tensor memory (tcgen05.ld 32x32b, one row per thread):
// either a plain store, or |X| -> ln(|X| + 1e-6) -> fftshift roll (layers/virtual_radar.py:131-133) written straight
// into the (N, n_fft, F) output, optionally saving Re/Im for the backward pass.
// The same kernel does the backward GEMMs (operands are addressed through element strides, so transposes cost nothing):
// dA = dC . Bt (gradient of the frames) and dBt = dC^T . A (gradient of the kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vr {

constexpr int GM = 128, GN = 128, GK = 32;          // CTA tile, K block per stage
#ifndef VR_GEMM_THREADS
#define VR_GEMM_THREADS 1024
#endif
constexpr int G_THREADS = VR_GEMM_THREADS;           // all warps stage the operands; warp w reads accumulator lanes 32 (w % 4)
constexpr int G_PARTS = G_THREADS / 128;             // ... and the (w / 4)-th part of the columns in the epilogue
constexpr int G_ACC = 4;                            // accumulators in tensor memory (G_ACC * GN = 512 columns)
constexpr int G_STAGE_BYTES = 2 * (GM + GN) * GK * 4;    // hi + lo of the A and B blocks: 64 KB
constexpr int G_SMEM_BYTES = 2 * G_STAGE_BYTES + 1024;

struct GemmParams {
    const float* A; long long sAm, sAk;             // A(m, k) = A[m * sAm + k * sAk]
    const float* B; long long sBn, sBk;             // B(n, k) = B[n * sBn + k * sBk]    (C = A . B^T)
    int M, N, K;
    int kb_per_split;                               // K blocks per gridDim.z slice (EPI 0 accumulates with atomics when gridDim.z > 1)
    float* C; long long ldc;                        // EPI 0: C[m * ldc + n]
    float* out; float* csave;                       // EPI 1: (sequences, n_fft, F) log-magnitude; optional raw C, column-major (N x ldc)
    int F, n_fft, nb;                               // frames per sequence; bins per column tile (re block | im block)
};

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46) (all >> 4),
    // version [46,48) = 1 on Blackwell, layout type [61,64) = 0 (no swizzle).  K-major, no swizzle: LBO = distance
    // between the two 16-byte K chunks of an instruction, SBO = distance between 8-row groups.
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// one row's 16 columns summed over the accumulators that were used (the first `nacc` big ones and the small one)
__device__ __forceinline__ void tmem_ld16_sum(uint32_t taddr, int nbig, float (&v)[16]) {
    tmem_ld16(taddr, v);
    float w[16];
    for (int a = 1; a < nbig; ++a) {
        tmem_ld16(taddr + a * GN, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
    }
    tmem_ld16(taddr + (G_ACC - 1) * GN, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += w[i];
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    } while (!ok);
}

__device__ __forceinline__ float tf32_rna(float a) {       // nearest TF32 (ties away): the split's lo part is then sign-symmetric
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
    return __uint_as_float(r);
}
// one operand block (ROWS x GK) from global memory into the canonical layout, split into hi / lo:
// element (r, k) at (k / 4) * (ROWS * 16) + r * 16 + (k % 4) * 4.  Two halves, so that the loads of the block after next
// are in flight while this one is split and multiplied (the loop was bound by the latency of these loads: ncu
// profiles/r02s, long-scoreboard 18 of 30 stall cycles per issue): g_load into registers, g_split from them.
// consecutive threads take consecutive ROWS of one 16-byte K chunk: conflict-free 16-byte shared-memory stores
// (the first version had consecutive threads on consecutive chunks of a row: 87 % of its shared-memory wavefronts
// were bank conflicts, ncu profiles/r02h); a thread's eight chunks of a row-major operand are one 128-byte line,
// and operands addressed with the row index contiguous (the backward GEMMs) load coalesced
// which 16-byte chunk (row r, K chunk kc) of the block a thread moves.  The shared-memory store is conflict-free when the
// eight lanes of a quarter warp hold eight consecutive rows of one chunk (row r sits at byte 16 r of its chunk plane).
//  - K contiguous in memory (row-major operand): a warp = 8 rows x 4 adjacent chunks, i.e. 8 lines with two whole
//    sectors each.  (32 rows x 1 chunk, the first mapping, made every load instruction touch 32 lines: the L1 tag stage
//    ran at 84 % of its peak and bounded the loop, ncu profiles/r02t.)
//  - rows contiguous in memory (the transposed operands of the backward GEMMs): a warp = 32 rows x 1 chunk, whose four
//    scalar loads are one line each.
template <int ROWS>
__device__ __forceinline__ void g_chunk(int c, bool k_contig, int& r, int& kc) {
    if (k_contig) {
        const int lane = c & 31, w = c >> 5;                  // warp-sized group w: rows 8 (w % (ROWS/8)) .., chunks 4 (w / (ROWS/8)) ..
        r = 8 * (w % (ROWS / 8)) + (lane & 7);
        kc = 4 * (w / (ROWS / 8)) + (lane >> 3);
    } else {
        r = c % ROWS;
        kc = c / ROWS;
    }
}
template <int ROWS>
struct GFrag { float4 v[ROWS * (GK / 4) / G_THREADS]; };
template <int ROWS>
__device__ __forceinline__ void g_load(GFrag<ROWS>& f, const float* __restrict__ src, long long s_row, long long s_k, int row0,
                                       int nrows, int k0, int K, int tid) {
    static_assert(ROWS * (GK / 4) % G_THREADS == 0, "whole chunks per thread");
#pragma unroll
    for (int i = 0; i < ROWS * (GK / 4) / G_THREADS; ++i) {
        int r, kc;
        g_chunk<ROWS>(tid + i * G_THREADS, s_k == 1, r, kc);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int row = row0 + r, k = k0 + 4 * kc;
        if (row < nrows) {
            const float* p = src + (long long)row * s_row + (long long)k * s_k;
            if (s_k == 1 && k + 3 < K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) v = __ldg(reinterpret_cast<const float4*>(p));
            else {
                if (k < K) v.x = __ldg(p);
                if (k + 1 < K) v.y = __ldg(p + s_k);
                if (k + 2 < K) v.z = __ldg(p + 2 * s_k);
                if (k + 3 < K) v.w = __ldg(p + 3 * s_k);
            }
        }
        f.v[i] = v;
    }
}
template <int ROWS>
__device__ __forceinline__ void g_split(unsigned char* hi, unsigned char* lo, const GFrag<ROWS>& f, bool k_contig, int tid) {
#pragma unroll
    for (int i = 0; i < ROWS * (GK / 4) / G_THREADS; ++i) {
        int r, kc;
        g_chunk<ROWS>(tid + i * G_THREADS, k_contig, r, kc);
        const float4 v = f.v[i];
        const float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);       // exact
        const int off = kc * (ROWS * 16) + r * 16;
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1) vr_gemm_tf32x3_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) unsigned char gsm[];
    __shared__ uint64_t bar_free[2], bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ntn = (p.N + GN - 1) / GN;                         // column tiles fastest: the CTAs that share an A tile run together
    const int tile_n = blockIdx.x % ntn, tile_m = blockIdx.x / ntn;    // (A once from HBM; it was read once per column tile, ncu r02s)
    const int m0 = tile_m * GM, n0 = tile_n * GN;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar_free[i])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar_done)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                              // one warp allocates (and later frees) the accumulators' 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"((uint32_t)(G_ACC * GN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6), A = B = TF32 [7,10) [10,13), both K-major,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
    const int KB_all = (p.K + GK - 1) / GK;
    const int kb_first = blockIdx.z * p.kb_per_split;
    const int KB = (KB_all - kb_first < p.kb_per_split) ? (KB_all - kb_first) : p.kb_per_split;     // this slice's K blocks
    GFrag<GM> fa[2];
    GFrag<GN> fb[2];
    const int k_first = kb_first * GK;
    g_load<GM>(fa[0], p.A, p.sAm, p.sAk, m0, p.M, k_first, p.K, tid);
    g_load<GN>(fb[0], p.B, p.sBn, p.sBk, n0, p.N, k_first, p.K, tid);
    if (KB > 1) {
        g_load<GM>(fa[1], p.A, p.sAm, p.sAk, m0, p.M, k_first + GK, p.K, tid);
        g_load<GN>(fb[1], p.B, p.sBn, p.sBk, n0, p.N, k_first + GK, p.K, tid);
    }
    auto k_block = [&](const int kb, GFrag<GM>& ra, GFrag<GN>& rb) {
        const int s = kb & 1;
        unsigned char* st = gsm + s * G_STAGE_BYTES;
        unsigned char* a_hi = st, *a_lo = st + GM * GK * 4, *b_hi = st + 2 * GM * GK * 4, *b_lo = b_hi + GN * GK * 4;
        if (kb >= 2) g_mbar_wait(&bar_free[s], (uint32_t)(((kb >> 1) - 1) & 1));    // the MMAs that read this stage have completed
        g_split<GM>(a_hi, a_lo, ra, p.sAk == 1, tid);
        g_split<GN>(b_hi, b_lo, rb, p.sBk == 1, tid);
        if (kb + 2 < KB) {                                        // the block after next: in flight over two barriers
            g_load<GM>(ra, p.A, p.sAm, p.sAk, m0, p.M, k_first + (kb + 2) * GK, p.K, tid);
            g_load<GN>(rb, p.B, p.sBn, p.sBk, n0, p.N, k_first + (kb + 2) * GK, p.K, tid);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core's reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = (uint32_t)__cvta_generic_to_shared(a_hi), al = (uint32_t)__cvta_generic_to_shared(a_lo);
            const uint32_t bh = (uint32_t)__cvta_generic_to_shared(b_hi), bl = (uint32_t)__cvta_generic_to_shared(b_lo);
#pragma unroll
            for (int ks = 0; ks < GK / 8; ++ks) {                 // K = 8 per instruction = two 16-byte chunks
                const uint32_t ao = ks * 2 * (GM * 16), bo = ks * 2 * (GN * 16);
                const uint64_t dah = umma_desc_kmajor(ah + ao, GM * 16, 128), dal = umma_desc_kmajor(al + ao, GM * 16, 128);
                const uint64_t dbh = umma_desc_kmajor(bh + bo, GN * 16, 128), dbl = umma_desc_kmajor(bl + bo, GN * 16, 128);
                const int step = kb * (GK / 8) + ks, big = step % (G_ACC - 1);
                umma_tf32(tmem + big * GN, dah, dbh, idesc, step >= G_ACC - 1 ? 1u : 0u);       // hi.hi: rotate over three accumulators
                umma_tf32(tmem + (G_ACC - 1) * GN, dal, dbh, idesc, step > 0 ? 1u : 0u);        // the small cross terms: the fourth
                umma_tf32(tmem + (G_ACC - 1) * GN, dah, dbl, idesc, 1u);
            }
            umma_commit(&bar_free[s]);                            // arrives when the MMAs issued so far have read shared memory
            if (kb == KB - 1) umma_commit(&bar_done);
        }
    };
    for (int kb = 0; kb < KB; kb += 2) {
        k_block(kb, fa[0], fb[0]);
        if (kb + 1 < KB) k_block(kb + 1, fa[1], fb[1]);
    }
    g_mbar_wait(&bar_done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: thread = accumulator row (TMEM lane 32 * (warp % 4) + lane); the warps of each group of four take one
    // part of the columns (bins) ----
    const int m = m0 + (tid & 127);
    const int part = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int steps = KB * (GK / 8), nbig = steps < G_ACC - 1 ? steps : G_ACC - 1;     // big accumulators that hold data
    if (EPI == 0) {
        for (int c0 = part * (GN / G_PARTS); c0 < (part + 1) * (GN / G_PARTS); c0 += 16) {
            float v[16];
            tmem_ld16_sum(trow + c0, nbig, v);
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + c0 + j < p.N) {
                        float* dst = p.C + (long long)m * p.ldc + n0 + c0 + j;
                        if (gridDim.z > 1) atomicAdd(dst, v[j]);        // split K: C was zeroed by the host
                        else *dst = v[j];
                    }
            }
        }
    } else {
        // column tile: [re of bins t*nb .. | im of the same bins]; out[(seq * n_fft + ((bin + n_fft/2) % n_fft)) * F + f]
        const int nb = p.nb, tile = tile_n;
        const int seq = m / p.F, f = m - seq * p.F;
        const int parts = (nb / 16 < G_PARTS) ? (nb / 16 > 0 ? nb / 16 : 1) : G_PARTS;      // 16-column loads: every part a multiple of 16 bins
        const int per = (nb / parts + 15) / 16 * 16;
        const int cb = part < parts ? part * per : 0, ce = part < parts ? (cb + per < nb ? cb + per : nb) : 0;
        for (int c0 = cb; c0 < ce; c0 += 16) {
            float re[16], im[16];
            tmem_ld16_sum(trow + c0, nbig, re);
            tmem_ld16_sum(trow + nb + c0, nbig, im);
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int bin = tile * nb + c0 + j;
                    if (c0 + j < nb && bin < p.n_fft) {
                        const float mag = sqrtf(fmaf(re[j], re[j], im[j] * im[j]));
                        const int row = (bin + p.n_fft / 2) % p.n_fft;
                        p.out[((long long)seq * p.n_fft + row) * p.F + f] = logf(mag + 1e-6f);
                        if (p.csave) {
                            p.csave[(long long)(n0 + c0 + j) * p.ldc + m] = re[j];          // column-major: lanes = rows, coalesced
                            p.csave[(long long)(n0 + nb + c0 + j) * p.ldc + m] = im[j];
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(G_ACC * GN)) : "memory");
}

// ---- small kernels around the GEMM -------------------------------------------------------------------------------
// frames: iq (S, T, 2) -> A (S*F, 2 n_fft): [I frame | Q frame], reflect-padded by n_fft/2 (nnAudio center=True)
__global__ void vr_stft_frames_kernel(const float* __restrict__ iq, float* __restrict__ A, long long S, int T, int F, int n_fft, int hop) {
    const long long total = S * F * (long long)n_fft;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % n_fft);
        const long long m = i / n_fft;
        const int f = (int)(m % F);
        const long long s = m / F;
        int t = f * hop - n_fft / 2 + k;
        t = t < 0 ? -t : t;
        t = t >= T ? 2 * (T - 1) - t : t;
        const float2 z = __ldg(reinterpret_cast<const float2*>(iq) + s * T + t);
        A[m * 2 * n_fft + k] = z.x;
        A[m * 2 * n_fft + n_fft + k] = z.y;
    }
}
// row of Bt that holds Re / Im of `bin` under the column-tile layout [re block | im block] of nb bins
__host__ __device__ inline int stft_row_re(int bin, int nb) { return (bin / nb) * 2 * nb + bin % nb; }
__host__ __device__ inline int stft_row_im(int bin, int nb) { return (bin / nb) * 2 * nb + nb + bin % nb; }
// Bt (2 n_fft, 2 n_fft) from stft.wsin / stft.wcos (n_fft, 1, n_fft)
__global__ void vr_stft_bt_kernel(const float* __restrict__ wsin, const float* __restrict__ wcos, float* __restrict__ Bt, int n_fft, int nb) {
    const int total = n_fft * n_fft;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int bin = i / n_fft, s = i - bin * n_fft;
        const float c = __ldg(wcos + i), sn = __ldg(wsin + i);
        float* re = Bt + (long long)stft_row_re(bin, nb) * 2 * n_fft;
        float* im = Bt + (long long)stft_row_im(bin, nb) * 2 * n_fft;
        re[s] = c; re[n_fft + s] = sn;
        im[s] = -sn; im[n_fft + s] = c;
    }
}
// backward of ln(|X| + 1e-6) and the roll: dC from grad_out and the saved Re / Im
__global__ void vr_stft_dc_kernel(const float* __restrict__ gout, const float* __restrict__ csave, float* __restrict__ dC,
                                  long long M, long long ldm, int F, int n_fft, int nb) {
    // C and dC are column-major (2 n_fft columns of ldm floats): threads run along the rows, every access coalesced
    const long long total = M * n_fft;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i % M;
        const int bin = (int)(i / M);
        const long long seq = m / F;
        const int f = (int)(m - seq * F);
        const int cr = stft_row_re(bin, nb), ci = stft_row_im(bin, nb);
        const float re = csave[cr * ldm + m], im = csave[ci * ldm + m];
        const float g = __ldg(gout + (seq * n_fft + (bin + n_fft / 2) % n_fft) * F + f);
        const float mag = sqrtf(fmaf(re, re, im * im));
        const float w = mag > 0.f ? g / ((mag + 1e-6f) * mag) : 0.f;
        dC[cr * ldm + m] = w * re;
        dC[ci * ldm + m] = w * im;
    }
}
// overlap-add of the frame gradients through the reflect padding: dA (S*F, 2 n_fft) -> grad_iq (S, T, 2), zeroed by the caller
__global__ void vr_stft_fold_kernel(const float* __restrict__ dA, float* __restrict__ giq, long long S, int T, int F, int n_fft, int hop) {
    const long long total = S * F * (long long)n_fft;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % n_fft);
        const long long m = i / n_fft;
        const int f = (int)(m % F);
        const long long s = m / F;
        int t = f * hop - n_fft / 2 + k;
        t = t < 0 ? -t : t;
        t = t >= T ? 2 * (T - 1) - t : t;
        atomicAdd(giq + (s * T + t) * 2, dA[m * 2 * n_fft + k]);
        atomicAdd(giq + (s * T + t) * 2 + 1, dA[m * 2 * n_fft + n_fft + k]);
    }
}
// dBt (2 n_fft, 2 n_fft) -> grad wsin / wcos
__global__ void vr_stft_dw_kernel(const float* __restrict__ dBt, float* __restrict__ gsin, float* __restrict__ gcos, int n_fft, int nb) {
    const int total = n_fft * n_fft;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int bin = i / n_fft, s = i - bin * n_fft;
        const float* re = dBt + (long long)stft_row_re(bin, nb) * 2 * n_fft;
        const float* im = dBt + (long long)stft_row_im(bin, nb) * 2 * n_fft;
        gcos[i] = re[s] + im[n_fft + s];
        gsin[i] = re[n_fft + s] - im[s];
    }
}
#endif  // __CUDACC__

}  // namespace vr
