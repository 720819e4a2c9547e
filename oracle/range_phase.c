/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the two rounding-critical steps of the
 * reference's VirtualRadar.forward:
 *
 *   distances = torch.norm(|src - radar_location|, dim=1)     layers/virtual_radar.py:96-99
 *   theta     = 4 * np.pi * distances / self.wavelength       layers/virtual_radar.py:119
 *
 * ATen's CPU f32 norm over the 3-element coordinate axis rounds in one of two ways depending on
 * the stride of that axis (re-derived on the running machine by
 * tests/test_oracle.py::test_aten_norm_matches_c_recipe):
 *
 *   mode 0 "seq": coordinate axis strided   ->  sqrt((a*a + b*b) + c*c), every op rounded, no FMA
 *   mode 1 "fma": coordinate axis innermost ->  sqrt(fma(c,c, fma(b,b, a*a)))
 *
 * theta is (f32(4*pi) * d) / lambda with each op rounded to f32 (python evaluates
 * `4*np.pi` in f64 first, the tensor op then casts that scalar to f32).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off so gcc never fuses a*a+b*b by itself).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <math.h>
#include <stdint.h>

/* s: (n,3) source-joint coordinates, AoS; loc: radar location (3); out_d, out_theta: (n). */
void vr_oracle_range_phase(const float* s, int64_t n, const float* loc, float wavelength,
                           int mode, float* out_d, float* out_theta) {
    const float four_pi = (float)12.566370614359172;
    for (int64_t i = 0; i < n; ++i) {
        /* |s - loc| : abs does not change the squares, kept for fidelity with :96 */
        volatile float a = fabsf(s[3 * i + 0] - loc[0]);
        volatile float b = fabsf(s[3 * i + 1] - loc[1]);
        volatile float c = fabsf(s[3 * i + 2] - loc[2]);
        float d2;
        if (mode == 0) {
            volatile float aa = a * a;
            volatile float bb = b * b;
            volatile float cc = c * c;
            volatile float ab = aa + bb;
            d2 = ab + cc;
        } else {
            volatile float aa = a * a;
            volatile float t = fmaf(b, b, aa);
            d2 = fmaf(c, c, t);
        }
        float d = sqrtf(d2);
        volatile float num = four_pi * d;
        out_d[i] = d;
        out_theta[i] = num / wavelength;
    }
}

/* Aspect cosine of a bone as the reference rounds it (layers/virtual_radar.py:101-105):
 *   A = loc - (S + D)/2,  B = D - S,  u = sum(A*B) / (norm(A) * norm(B) + 1e-6)
 * torch.sum over the 3 coordinates = (p0 + p1) + p2 of individually rounded products (both
 * layouts); the two norms use the layout's mode like the range above; sqrt, multiply, add and
 * divide are single IEEE f32 operations.  s, d: (n,3) AoS bone end points.  Also returns |B|
 * (the bone length entering the mean at :110-112).                                            */
static float norm3(float a, float b, float c, int mode) {
    if (mode == 0) {
        volatile float aa = a * a; volatile float bb = b * b; volatile float cc = c * c;
        volatile float ab = aa + bb;
        volatile float t = ab + cc;
        return sqrtf(t);
    } else {
        volatile float aa = a * a;
        volatile float t = fmaf(b, b, aa);
        volatile float t2 = fmaf(c, c, t);
        return sqrtf(t2);
    }
}

void vr_oracle_aspect_cosine(const float* s, const float* d, int64_t n, const float* loc, int mode,
                             float* out_u, float* out_len) {
    const float eps = (float)1e-6;
    for (int64_t i = 0; i < n; ++i) {
        float A[3], B[3];
        for (int c = 0; c < 3; ++c) {
            volatile float sum = s[3 * i + c] + d[3 * i + c];
            volatile float mid = sum / 2.0f;
            A[c] = loc[c] - mid;
            B[c] = d[3 * i + c] - s[3 * i + c];
        }
        volatile float p0 = A[0] * B[0]; volatile float p1 = A[1] * B[1]; volatile float p2 = A[2] * B[2];
        volatile float p01 = p0 + p1;
        volatile float dot = p01 + p2;
        float na = norm3(A[0], A[1], A[2], mode), nb = norm3(B[0], B[1], B[2], mode);
        volatile float q = na * nb;
        volatile float qe = q + eps;
        out_u[i] = dot / qe;
        out_len[i] = nb;
    }
}

int vr_oracle_abi(void) { return 2; }
