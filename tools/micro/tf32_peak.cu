// Legacy-path tensor throughput on B200: mma.sync m16n8k8 tf32 and m16n8k16 bf16 (register-only loops).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_tf32(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_bf16(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}
int main() {
    float* d; cudaMalloc(&d, 256);
    const int iters = 20000, blocks = 148 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (which == 0) k_tf32<<<blocks, 256>>>(d, iters); else k_bf16<<<blocks, 256>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
        }
        double flop = (double)blocks * 8 * iters * 8 * (which == 0 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16);
        printf("%s mma.sync: %.3f ms  %.1f TFLOP/s\n", which == 0 ? "tf32 m16n8k8" : "bf16 m16n8k16", best, flop / (best * 1e-3) / 1e12);
    }
    return 0;
}
