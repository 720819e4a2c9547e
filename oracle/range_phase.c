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

int vr_oracle_abi(void) { return 1; }
