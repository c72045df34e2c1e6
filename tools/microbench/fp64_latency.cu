// Developer microbenchmark: dependent-chain latency of FP64 ops on the target GPU (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { x = __fma_rn(x, y, 1e-9); x = __fma_rn(x, y, 1e-9); x = __fma_rn(x, y, 1e-9); x = __fma_rn(x, y, 1e-9); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { x = __dadd_rn(x, y); x = __dadd_rn(x, y); x = __dadd_rn(x, y); x = __dadd_rn(x, y); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { x = __dmul_rn(x, y); x = __dmul_rn(x, y); x = __dmul_rn(x, y); x = __dmul_rn(x, y); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // division chain
    x = fabs(x) + 1.5;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { x = 3.0 / x + 1.0; x = 3.0 / x + 1.0; x = 3.0 / x + 1.0; x = 3.0 / x + 1.0; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // sqrt chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { x = sqrt(x) + 1.0; x = sqrt(x) + 1.0; x = sqrt(x) + 1.0; x = sqrt(x) + 1.0; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 2 independent division chains (ILP)
    double z = x + 0.25;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { x = 3.0 / x + 1.0; z = 3.0 / z + 1.0; x = 3.0 / x + 1.0; z = 3.0 / z + 1.0; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // FFMA chain for reference
    float f = (float)x, g = (float)y;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { f = __fmaf_rn(f, g, 1e-9f); f = __fmaf_rn(f, g, 1e-9f); f = __fmaf_rn(f, g, 1e-9f); f = __fmaf_rn(f, g, 1e-9f); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = x + z + f;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 0.999999, 1.0000001, n); cudaDeviceSynchronize(); }
    const char* names[] = {"DFMA", "DADD", "DMUL", "DDIV(+add)", "DSQRT(+add)", "2x DDIV interleaved (per pair)", "FFMA"};
    for (int i = 0; i < 7; ++i) printf("%-32s %8.1f cycles per dependent op\n", names[i], (double)cyc[i] / (4.0 * n) * (i == 5 ? 2 : 1));
    return 0;
}
