// Are the FP64 tensor pipe (DMMA) and the FP64 FMA pipe independent on B200?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NMMA, int NFMA>
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double c[8][2], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) dmma(c[i & 7][0], c[i & 7][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i & 15]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}
template <int NMMA, int NFMA>
void run(const char* name) {
    double* d; cudaMalloc(&d, 256);
    const int iters = 20000, blocks = 148 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<NMMA, NFMA><<<blocks, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    double warps = (double)blocks * 8 * iters;
    double tf_mma = warps * NMMA * 512.0 / (best * 1e-3) / 1e12, tf_fma = warps * NFMA * 64.0 / (best * 1e-3) / 1e12;
    printf("%-28s %8.3f ms  DMMA %6.2f TF/s  DFMA %6.2f TF/s  total %6.2f TF/s\n", name, best, tf_mma, tf_fma, tf_mma + tf_fma);
    cudaFree(d);
}
int main() {
    run<8, 0>("DMMA only (8/iter)");
    run<0, 64>("DFMA only (64/iter)");
    run<8, 16>("8 DMMA + 16 DFMA");
    run<8, 32>("8 DMMA + 32 DFMA");
    run<8, 64>("8 DMMA + 64 DFMA");
    run<8, 128>("8 DMMA + 128 DFMA");
    return 0;
}
