// microbenchmark: DFMA and F2F.F32.F64 issue rates per SM on this GPU (build: nvcc -arch=sm_100a -O3)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.000001, c = 1e-9, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); }
    if (a + d + e + f == 1234.5) out[0] = a;
}
__global__ void f2f(float* out, int iters) {
    double a = threadIdx.x * 1e-3, d = a + 1, e = a + 2, f = a + 3;
    float s = 0.f;
    for (int i = 0; i < iters; ++i) { s += (float)a + (float)d + (float)e + (float)f; a += 1.0; d += 1.0; e += 1.0; f += 1.0; }
    if (s == 1234.5f) out[0] = s;
}
__global__ void ffma(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = 1.000001f, c = 1e-9f, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < iters; ++i) { a = fmaf(a, b, c); d = fmaf(d, b, c); e = fmaf(e, b, c); f = fmaf(f, b, c); }
    if (a + d + e + f == 1234.5f) out[0] = a;
}
int main() {
    double* o; cudaMalloc(&o, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k = 0; k < 3; ++k) {
        float ms;
        dfma<<<sms * 2, 1024>>>(o, 100);
        cudaEventRecord(e0); dfma<<<sms * 2, 1024>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)sms * 2 * 1024 * iters * 4;
        printf("DFMA: %.2f T/s  (%.1f per SM per clk at 1.9 GHz)\n", n / ms / 1e9, n / ms / 1e6 / sms / 1.9e3);
        cudaEventRecord(e0); f2f<<<sms * 2, 1024>>>((float*)o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("F2F+DADD (4+4 per iter): %.2f T iter-lanes/s (%.1f F2F per SM per clk)\n", n / ms / 1e9, n / ms / 1e6 / sms / 1.9e3);
        cudaEventRecord(e0); ffma<<<sms * 2, 1024>>>((float*)o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA: %.2f T/s  (%.1f per SM per clk)\n", n / ms / 1e9, n / ms / 1e6 / sms / 1.9e3);
    }
    return 0;
}
