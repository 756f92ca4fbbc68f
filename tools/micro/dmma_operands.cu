// Does the per-warp DMMA cadence depend on where the operands come from?  One warp per scheduler (4 warps per SM), 12
// independent accumulator chains; operands (a) the same register pair for every DMMA, (b) rotating through 8 + 8 distinct
// registers (the pattern of a real update / GEMM loop), (c) as (b) with the B operands re-loaded from shared memory every
// 12 DMMAs.   nvcc -arch=sm_100a -O3 -o dmma_operands dmma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, int iters) {
    __shared__ double sb[8][32];
    double c[12][2];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i][0] = c[i][1] = 0.0;
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0 + i * 1e-4; sb[i][threadIdx.x & 31] = b[i]; }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = sb[(k4 + j) & 7][threadIdx.x & 31];
            }
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                if (MODE == 0) dmma(c[j][0], c[j][1], a[0], b[0]);
                else dmma(c[j][0], c[j][1], a[k4], b[j & 3]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int MODE>
void run(double* d, int warps, const char* what) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<148, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double n = (double)iters * 96;
    printf("%-44s warps/SM %2d: %.1f clk per DMMA per warp, %.3f DMMA/clk/SM (peak 0.25)\n", what, warps, best * 1e-3 * 1.965e9 / n, n * warps / (best * 1e-3 * 1.965e9));
}
int main() {
    double* d; cudaMalloc(&d, 256);
    for (int w : {4, 8, 12}) {
        run<0>(d, w, "same operand registers");
        run<1>(d, w, "rotating operand registers");
        run<2>(d, w, "rotating + B re-loaded from shared memory");
    }
    return 0;
}
