// Dependent-chain latency of the FP64 instructions the Jacobi eigensolve is made of (one warp, one SM):
// DFMA, DMUL, rsqrt(double), and an LDS.128 -> DFMA -> STS.128 round trip.   nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, int iters) {
    __shared__ double2 sm[64];
    double x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
    sm[threadIdx.x] = make_double2(x, y);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) x = x * y;
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) x = rsqrt(x + 2.0);
    long long t3 = clock64();
    for (int i = 0; i < iters; ++i) {
        double2 v = sm[(threadIdx.x + i) & 63];
        v.x = fma(v.x, y, x);
        sm[(threadIdx.x + i) & 63] = v;
        __syncwarp();
        x = v.x * 1e-3;
    }
    long long t4 = clock64();
    float f = (float)x;
    for (int i = 0; i < iters; ++i) f = fmaf(f, 0.999f, 1e-6f);
    long long t5 = clock64();
    for (int i = 0; i < iters; ++i) { __syncthreads(); }
    long long t6 = clock64();
    if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t5 - t4; clk[5] = t6 - t5; }
    out[threadIdx.x] = x + f;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64);
    const int iters = 4096;
    for (int threads : {32, 128}) {
        k<<<1, threads>>>(d, c, iters); cudaDeviceSynchronize();
        k<<<1, threads>>>(d, c, iters); cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%3d threads: DFMA %.1f clk, DMUL %.1f clk, rsqrt(double) %.1f clk, LDS.128->DFMA->STS.128->DMUL %.1f clk, FFMA %.1f clk, __syncthreads %.1f clk (dependent chains)\n",
               threads, (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, (double)h[5] / iters);
    }
    return 0;
}
