// Per-warp DMMA issue limits on B200: throughput of mma.sync.m8n8k4.f64 as a function of resident warps per SM and of
// the number of independent accumulator chains per warp (register-only loops).  Explains how many tensor-phase warps
// the Jacobi dataflow kernel needs per SM.   nvcc -arch=sm_100a -O3 -o dmma_warp dmma_warp.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
void run(double* d, int warps_per_sm) {
    const int iters = 40000 / NACC;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<NACC><<<148, warps_per_sm * 32>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double n = (double)iters * NACC;                     // DMMAs per warp
    const double clk = best * 1e-3 * 1.965e9;
    printf("warps/SM %2d  chains %2d : %.1f clk per DMMA per warp, %.2f DMMA/clk/SM (peak 0.25), %.1f TFLOP/s\n", warps_per_sm, NACC,
           clk / n, n * warps_per_sm / clk, 148.0 * warps_per_sm * n * 512 / (best * 1e-3) / 1e12);
}
int main() {
    double* d; cudaMalloc(&d, 256);
    for (int w : {1, 2, 4, 8, 16}) { run<1>(d, w); run<3>(d, w); run<6>(d, w); run<12>(d, w); run<24>(d, w); }
    return 0;
}
