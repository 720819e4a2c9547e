// microbenchmark: dependent-issue latency of DFMA, F2F.F32.F64 and I2F.F64 on this GPU (one warp per SM, clock64)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(long long* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    float f = (float)a;
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) { f = (float)((double)f * b); }       // F2F.F64.F32 + DMUL + F2F.F32.F64 chain
    long long t3 = clock64();
    float g = f;
    for (int i = 0; i < iters; ++i) g = fmaf(g, 1.000001f, 1e-9f);
    long long t4 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t3 - t2; out[2] = t4 - t3; }
    if (a + f + g == 1234.5) out[3] = 1;
}
__global__ void lat_loaded(long long* out, int iters) {      // same DFMA chain with 32 warps per SM resident (1024 threads)
    double a = threadIdx.x * 1e-3, b = 1.000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[4] = t1 - t0;
    if (a == 1234.5) out[3] = 1;
}
int main() {
    long long* o; cudaMalloc(&o, 64); cudaMemset(o, 0, 64);
    int iters = 10000;
    lat<<<148, 32>>>(o, iters); lat<<<148, 32>>>(o, iters);
    lat_loaded<<<148, 1024>>>(o, iters);
    long long h[8]; cudaMemcpy(h, o, 64, cudaMemcpyDeviceToHost);
    printf("dependent DFMA: %.1f cycles;  F2F+DMUL+F2F chain: %.1f cycles;  dependent FFMA: %.1f cycles\n", h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters);
    printf("dependent DFMA with 32 warps per SM all doing the same: %.1f cycles per step per warp\n", h[4] / (double)iters);
    return 0;
}
