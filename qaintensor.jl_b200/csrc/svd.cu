// K5/K6: truncated SVD on the device (replaces LinearAlgebra.svd -> LAPACK zgesdd at
// src/svd.jl:26-27, src/switch.jl:39, src/mps.jl:63,73, src/mpo.jl:53,62,
// src/decompose.jl:26,38 and the tail-norm rule of src/svd.jl:29-33).
//
// Algorithm: blocked one-sided (Hestenes) Jacobi, batched over independent matrices.
//   A (m x n, m >= n) is orthogonalised in place, V accumulates the column rotations.
//   Columns are grouped in blocks of 16; a round-robin (circle method) schedule pairs the blocks.  A visit of a block
//   pair is
//     1. Gram   G = P^H P of the 32-column panel P on the FP64 tensor pipe (DMMA); in the linear-convergence sweeps only
//               the 16 x 16 cross block, the diagonal blocks come from a cache the eigensolve keeps current,
//     2. eig    two-sided cyclic Jacobi sweep on the 32x32 Hermitian G (parallel ordering, de Rijk-style sorting),
//               accumulating W; convergence stamps per block pair,
//     3. update P <- P W for the A panel and the matching V panel, again on DMMA.
//   Two schedulers:
//     * batches with >= 2 * 148 block pairs per round: ONE dataflow kernel per sweep (jacobi_flow_kernel) -- persistent
//       CTAs draw (matrix, round, pair) tasks from an atomic counter, run all three phases of a visit with G and W in
//       shared memory, and synchronise through per-block version flags instead of round barriers;
//     * smaller batches / single matrices: three kernels per round (Gram partials with the rows split over S CTAs,
//       one eigensolve CTA per pair, update with the rows split over SU CTAs), size-sorted sub-batches on concurrent
//       streams, or -- a single sub-batch -- chained by programmatic dependent launch.
//   Sweeps repeat until no visit applied a rotation.
//   Finalisation: sigma_j = ||a_j||, stable descending sort, U = A/sigma, Vh = V^H.
//   Truncation (K6): k = n - r* + 1 with r* the first r whose reverse-cumulated tail
//   sqrt(s_n^2 + ... + s_{n-r+1}^2) exceeds er (strict), then k <- min(k, maxdim).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "qtn_internal.h"

namespace qtn {
cudaStream_t stream();
void count_launch(int64_t n);
int permutedims_device(const void* in, int rank, const int64_t* dims, const int32_t* perm, void* out);
int launch_gemm(GemmArgs& g, int variant, int split_k, cudaStream_t st);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(QTN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

constexpr int JB = 16;        // column block width
constexpr int JP = 2 * JB;    // panel width (columns per CTA)
constexpr int JPITCH = JP + 2;
#ifndef QTN_UPDATE_THREADS
#define QTN_UPDATE_THREADS 128
#endif
constexpr int JTHREADS = QTN_UPDATE_THREADS;  // update kernel: small CTAs pack next to the eigensolve CTAs of other streams
constexpr int kInnerSweeps = 1;  // one eigen-sweep per Gram visit measured fastest (1: 108 ms, 2: 153, 3: 165, 6: 187 ms for 1024^2)

struct SvdProblem {
    double2* A;   // m x n (lda = m), overwritten with U * diag(S) (unsorted)
    double2* V;   // n x n
    int m, n;
    int nblocks;  // ceil(n / JB), rounded up to even
    int* last_mod;  // [nblocks]  launch stamp of the last rotation that touched a column block
    int* last_ok;   // [nblocks * nblocks] launch stamp at which a block pair was last found converged
    double2* D;     // [nblocks][16 x 16] Gram of each column block (row-major), kept current by the eigensolve kernel
    int* ver;       // [nblocks]  dataflow kernel: number of rounds of the current sweep this block has been through
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- optional CTA-level tracer (QTN_JACOBI_TRACE=<file>): every working CTA of the round kernels of ONE sweep
// records (kind, SM id, start, end) from %globaltimer -- the concurrent timeline of the sub-batch streams, which ncu
// (it serialises kernels) cannot show.  No cost when off (one predictable branch on a __device__ pointer).
__device__ unsigned long long* d_trace_buf = nullptr;
__device__ unsigned int d_trace_cap = 0;
__device__ unsigned int d_trace_cnt = 0;
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_end(int kind, unsigned long long t0) {
    if (d_trace_buf == nullptr || threadIdx.x != 0) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned i = atomicAdd(&d_trace_cnt, 1u);
    if (i < d_trace_cap) {
        d_trace_buf[3 * (size_t)i] = (unsigned long long)kind | ((unsigned long long)smid << 8) | ((unsigned long long)blockIdx.y << 24);
        d_trace_buf[3 * (size_t)i + 1] = t0;
        d_trace_buf[3 * (size_t)i + 2] = trace_now();
    }
}

// Programmatic dependent launch (single-matrix rounds are three short dependent kernels: the launch latency between them
// is a quarter of a round).  A kernel launched with the attribute may start while its predecessor drains; it must not
// touch the predecessor's results before pdl_wait(), and it lets ITS successor in with pdl_trigger() -- at the END of a
// CTA's work: triggered at the top, the successor's CTAs sat spinning on SM slots that the predecessor's later waves
// needed (4 x 512^2: 1.97 -> 2.64 ms per sweep).  Both are no-ops in launches without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// column index of panel slot c (0..31) or -1 when outside the matrix
__device__ __forceinline__ int panel_col(int c, int bi, int bj, int n) {
    const int col = (c < JB ? bi * JB + c : bj * JB + (c - JB));
    return col < n ? col : -1;
}

// =============================================================================================================
// The round kernels.  The three phases are separate kernels so that the tensor-pipe work (Gram, update) streams at
// GEMM efficiency and the latency-bound 32 x 32 eigensolves of ALL block pairs of a round run side by side:
//   jacobi_gram_kernel    partial Grams of every active (matrix, pair), rows split over S CTAs; operands go
//                         global -> registers as DMMA fragments (each element is read exactly once), no shared-memory
//                         staging and no barrier in the main loop; fixed-order cross-warp tree -> deterministic sums
//   jacobi_eig_kernel     one small CTA per (matrix, pair): G = sum of the S partials (fixed order), Hermitian Jacobi
//                         sweep, W and the pair's "apply" flag to global memory, convergence stamps
//   jacobi_update_kernel  P <- P W for the A panel and the V panel, 8-row groups per warp, W in shared memory
// =============================================================================================================
constexpr int GP_ELEMS = 10 * 64;   // upper block triangle of the 4 x 4 grid of 8 x 8 blocks of a 32 x 32 Gram
constexpr int GRAM_THREADS = 128;
#ifndef GRAM_MAXREG
#define GRAM_MAXREG 224  // 2 CTAs x 128 x 224 = 56 K registers: leaves exactly one eigensolve CTA (128 x 64) of another stream
#endif

// block pair of `pair` in `round` (circle method); false when this (round, pair) has nothing to do for the problem
__device__ __forceinline__ bool round_pair(const SvdProblem& pr, int round, int pair, int& bi, int& bj) {
    const int nb = pr.nblocks;
    if (nb < 2 || pair >= nb / 2 || round >= nb - 1) return false;
    if (pair == 0) { bi = nb - 1; bj = round; }
    else { bi = (round + pair) % (nb - 1); bj = (round - pair + (nb - 1)) % (nb - 1); }
    if (bi > bj) { const int t = bi; bi = bj; bj = t; }
    return bi * JB < pr.n;  // false: padding block only
}
// a pair verified converged stays converged until one of its two blocks is rotated again
__device__ __forceinline__ bool pair_idle(const SvdProblem& pr, int bi, int bj) {
    return max(pr.last_mod[bi], pr.last_mod[bj]) < pr.last_ok[bi * pr.nblocks + bj];
}

__global__ void __maxnreg__(GRAM_MAXREG)
jacobi_gram_kernel(const SvdProblem* __restrict__ probs, int round, const int* __restrict__ rotated, int S, int maxpairs,
                   double2* __restrict__ Gpart) {
    pdl_wait();
    const int b = blockIdx.y, pair = blockIdx.x / S, split = blockIdx.x - pair * S;
    if (rotated[b] < 0) return;  // matrix converged in an earlier sweep
    const SvdProblem pr = probs[b];
    int bi, bj;
    if (!round_pair(pr, round, pair, bi, bj) || pair_idle(pr, bi, bj)) return;
    const unsigned long long trace_t0 = d_trace_buf ? trace_now() : 0ull;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m = pr.m;
    // lane (g, t) holds P[row 4s + t][column 8i + g] of every 16-row group: the A fragment (as conj) and the B fragment
    // of mma.m8n8k4 coincide, so one 16-byte load per (s, i) feeds all ten blocks
    const double2* cp[4];
    bool cv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = panel_col(i * 8 + g, bi, bj, pr.n);
        cv[i] = col >= 0;
        cp[i] = pr.A + (size_t)(col >= 0 ? col : 0) * m;
    }
    const int ngroups = (m + 15) >> 4, gper = (ngroups + S - 1) / S;
    const int g0 = split * gper, g1 = min(ngroups, g0 + gper);
    double gr[10][2], gi[10][2];
#pragma unroll
    for (int k = 0; k < 10; ++k) gr[k][0] = gr[k][1] = gi[k][0] = gi[k][1] = 0.0;
    double2 v[4][4], w[4][4];
    auto load = [&](double2 (&dst)[4][4], int grp) {
        const int r = grp * 16 + t;
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = r + 4 * s;
                dst[s][i] = (cv[i] && row < m) ? __ldcg(cp[i] + row) : make_double2(0.0, 0.0);
            }
    };
    auto compute = [&](const double2 (&f)[4][4]) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int k = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = i; j < 4; ++j, ++k) {
                    dmma(gr[k][0], gr[k][1], f[s][i].x, f[s][j].x);
                    dmma(gr[k][0], gr[k][1], f[s][i].y, f[s][j].y);
                    dmma(gi[k][0], gi[k][1], f[s][i].x, f[s][j].y);
                    dmma(gi[k][0], gi[k][1], -f[s][i].y, f[s][j].x);
                }
        }
    };
    int grp = g0 + warp;
    constexpr int NW = GRAM_THREADS / 32;
    if (grp < g1) load(v, grp);
    while (grp < g1) {  // register double buffering: the next group's 16 loads are in flight during the 160 DMMAs
        if (grp + NW < g1) load(w, grp + NW);
        compute(v);
        grp += NW;
        if (grp >= g1) break;
        if (grp + NW < g1) load(v, grp + NW);
        compute(w);
        grp += NW;
    }
    // fixed-order tree over the 4 warps (deterministic, unlike atomics): 3+2 -> 1+0, then 1 -> 0
    __shared__ double red[2][40][32];
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {
        const int half = stage == 0 ? 2 : 1;  // warps [half, 2*half) hand over to warps [0, half)
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                red[warp - half][4 * k + 0][lane] = gr[k][0];
                red[warp - half][4 * k + 1][lane] = gr[k][1];
                red[warp - half][4 * k + 2][lane] = gi[k][0];
                red[warp - half][4 * k + 3][lane] = gi[k][1];
            }
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                gr[k][0] += red[warp][4 * k + 0][lane];
                gr[k][1] += red[warp][4 * k + 1][lane];
                gi[k][0] += red[warp][4 * k + 2][lane];
                gi[k][1] += red[warp][4 * k + 3][lane];
            }
        }
        __syncthreads();
    }
    if (warp == 0) {  // block k as 8 x 8 row-major: lane (g, t) owns columns 2t, 2t+1 of row g -> 32 contiguous bytes
        double2* out = Gpart + ((size_t)((size_t)b * maxpairs + pair) * S + split) * GP_ELEMS;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            out[k * 64 + g * 8 + 2 * t] = make_double2(gr[k][0], gi[k][0]);
            out[k * 64 + g * 8 + 2 * t + 1] = make_double2(gr[k][1], gi[k][1]);
        }
    }
    pdl_trigger();   // this CTA's results are on their way: the dependent grid may fill the SMs this grid's tail frees
    trace_end(0, trace_t0);
}

// Cross-only Gram (linear-convergence phase): only the 16 x 16 block A_I^H A_J is computed; the two diagonal blocks
// A_I^H A_I, A_J^H A_J come from the per-block cache SvdProblem::D, which the eigensolve kernel transforms along with
// every rotation (D' = W^H G W restricted to the block: exact up to rounding, refreshed by a full Gram in round 0 of
// every sweep and in all polishing sweeps).  4 of the 10 blocks remain, and conj(a) b takes 3 real products:
//   M1 = x_i x_j, M2 = y_i y_j, M3 = (x_i + y_i)(x_j - y_j):  re = M1 + M2, im = M1 - M2 - M3
// -> 12 DMMAs per 4 rows instead of 40; the kernel is HBM-bound (every panel element is still read once).
__global__ void __launch_bounds__(GRAM_THREADS, 2)
jacobi_gram_cross_kernel(const SvdProblem* __restrict__ probs, int round, const int* __restrict__ rotated, int S, int maxpairs,
                         double2* __restrict__ Gpart) {
    pdl_wait();
    const int b = blockIdx.y, pair = blockIdx.x / S, split = blockIdx.x - pair * S;
    if (rotated[b] < 0) return;
    const SvdProblem pr = probs[b];
    int bi, bj;
    if (!round_pair(pr, round, pair, bi, bj) || pair_idle(pr, bi, bj)) return;
    const unsigned long long trace_t0 = d_trace_buf ? trace_now() : 0ull;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m = pr.m;
    const double2* cp[4];
    bool cv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = panel_col(i * 8 + g, bi, bj, pr.n);
        cv[i] = col >= 0;
        cp[i] = pr.A + (size_t)(col >= 0 ? col : 0) * m;
    }
    const int ngroups = (m + 15) >> 4, gper = (ngroups + S - 1) / S;
    const int g0 = split * gper, g1 = min(ngroups, g0 + gper);
    double m1[4][2], m2[4][2], m3[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) m1[k][0] = m1[k][1] = m2[k][0] = m2[k][1] = m3[k][0] = m3[k][1] = 0.0;
    double2 v[4][4], w[4][4];
    auto load = [&](double2 (&dst)[4][4], int grp) {
        const int r = grp * 16 + t;
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = r + 4 * s;
                dst[s][i] = (cv[i] && row < m) ? __ldcg(cp[i] + row) : make_double2(0.0, 0.0);
            }
    };
    auto compute = [&](const double2 (&f)[4][4]) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double sa0 = f[s][0].x + f[s][0].y, sa1 = f[s][1].x + f[s][1].y;
            const double db2 = f[s][2].x - f[s][2].y, db3 = f[s][3].x - f[s][3].y;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = 2 * i + j;
                    dmma(m1[k][0], m1[k][1], f[s][i].x, f[s][2 + j].x);
                    dmma(m2[k][0], m2[k][1], f[s][i].y, f[s][2 + j].y);
                    dmma(m3[k][0], m3[k][1], i ? sa1 : sa0, j ? db3 : db2);
                }
        }
    };
    int grp = g0 + warp;
    constexpr int NW = GRAM_THREADS / 32;
    if (grp < g1) load(v, grp);
    while (grp < g1) {
        if (grp + NW < g1) load(w, grp + NW);
        compute(v);
        grp += NW;
        if (grp >= g1) break;
        if (grp + NW < g1) load(v, grp + NW);
        compute(w);
        grp += NW;
    }
    double gr[4][2], gi[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < 2; ++q) { gr[k][q] = m1[k][q] + m2[k][q]; gi[k][q] = m1[k][q] - m2[k][q] - m3[k][q]; }
    __shared__ double red[2][16][32];
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {  // fixed-order tree over the 4 warps, as in the full kernel
        const int half = stage == 0 ? 2 : 1;
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                red[warp - half][4 * k + 0][lane] = gr[k][0];
                red[warp - half][4 * k + 1][lane] = gr[k][1];
                red[warp - half][4 * k + 2][lane] = gi[k][0];
                red[warp - half][4 * k + 3][lane] = gi[k][1];
            }
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                gr[k][0] += red[warp][4 * k + 0][lane];
                gr[k][1] += red[warp][4 * k + 1][lane];
                gi[k][0] += red[warp][4 * k + 2][lane];
                gi[k][1] += red[warp][4 * k + 3][lane];
            }
        }
        __syncthreads();
    }
    if (warp == 0) {  // block k = 2 i + (j - 2) as 8 x 8 row-major in the first 256 elements of the partial
        double2* out = Gpart + ((size_t)((size_t)b * maxpairs + pair) * S + split) * GP_ELEMS;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            out[k * 64 + g * 8 + 2 * t] = make_double2(gr[k][0], gi[k][0]);
            out[k * 64 + g * 8 + 2 * t + 1] = make_double2(gr[k][1], gi[k][1]);
        }
    }
    pdl_trigger();
    trace_end(1, trace_t0);
}

// Rotation of the Hermitian 2x2 pivot [[alpha, g], [conj(g), beta]]: J = [[c, s e^{i phi}], [-s e^{-i phi}, c]].
struct JRot { double c, s; double2 ph; };

// ET threads per CTA.  A tournament step is ~400 instructions per warp at ET = 128 (ncu, uncontended: 6.8 cycles per issued
// instruction: FP64 chains, shared-memory round trips, two barriers), i.e. latency-bound with one warp per scheduler.
// Launches that cannot fill the machine anyway (single matrices: cfg 5's truncating sweep, switch!, contract_svd) run
// ET = 256 -- one 2x2 block of G and two rows of W per thread, two warps per scheduler: 5.12 -> 4.63 ms per sweep of a
// 1536 x 1024 matrix (ET = 512 with the blocks split over two threads needs a third barrier per step: 4.89).  Batched
// launches keep ET = 128: small CTAs that co-reside with the DMMA kernels of the other streams.
template <int ET>
__global__ void __launch_bounds__(ET, ET == 128 ? 8 : 4)
jacobi_eig_kernel(const SvdProblem* __restrict__ probs, int round, double tol, int* __restrict__ rotated,
                  const double* __restrict__ fro2, int inner_sweeps, int* __restrict__ stat, int stamp, int S, int maxpairs,
                  const double2* __restrict__ Gpart, double2* __restrict__ Wbuf, int* __restrict__ pflag, int cross,
                  int* __restrict__ nactive, int xrot) {
    pdl_wait();
    const int b = blockIdx.y, pair = blockIdx.x, tid = threadIdx.x;
    int* flag = pflag + (size_t)b * maxpairs + pair;
    const SvdProblem pr = probs[b];
    int bi = 0, bj = 0;
    const bool live = rotated[b] >= 0 && round_pair(pr, round, pair, bi, bj);
    if (!live || pair_idle(pr, bi, bj)) {
        if (tid == 0) {
            *flag = 0;
            if (live && stat) atomicAdd(&stat[0], 1);
        }
        return;
    }
    const unsigned long long trace_t0 = d_trace_buf ? trace_now() : 0ull;
    const int nb = pr.nblocks;
    // deflation: a column below 1e-15 * ||A||_F is rounding noise (LAPACK resolves nothing there either); rotating it
    // against a large, exactly parallel column would only shrink it by eps per sweep until it underflows
    // (rank-deficient product states hit exactly this)
    const double dthr = 1e-30 * fro2[b];
    __shared__ double2 G[JP * JPITCH], W[JP * JPITCH];
    __shared__ double2 rot[JB], rph[JB];
    __shared__ int s_pq[2 * JB];
    __shared__ int s_any, s_sweep_any, s_rot;
    __shared__ int s_cols[JP];
    if (tid < JP) s_cols[tid] = panel_col(tid, bi, bj, pr.n);
    if (tid == 0) { s_any = 0; s_rot = 0; }
    for (int i = tid; i < JP * JPITCH; i += ET) W[i] = make_double2(0, 0);
    double2* DI = pr.D + (size_t)bi * (JB * JB);
    double2* DJ = pr.D + (size_t)bj * (JB * JB);
    {
        const double2* gp = Gpart + (size_t)((size_t)b * maxpairs + pair) * S * GP_ELEMS;
        const int nel = cross ? 4 * 64 : GP_ELEMS;
        for (int e = tid; e < nel; e += ET) {
            double2 acc = make_double2(0, 0);
            for (int s = 0; s < S; ++s) {  // fixed order: deterministic
                const double2 v = __ldcg(gp + (size_t)s * GP_ELEMS + e);
                acc.x += v.x;
                acc.y += v.y;
            }
            const int k = e >> 6, r8 = (e >> 3) & 7, c8 = e & 7;
            int i, j;
            if (cross) { i = k >> 1; j = 2 + (k & 1); }  // the four blocks of A_I^H A_J
            else {
                // k -> (i, j), j >= i: rows of the upper block triangle hold 4, 3, 2, 1 blocks
                i = k < 4 ? 0 : (k < 7 ? 1 : (k < 9 ? 2 : 3));
                j = k - (i == 0 ? 0 : (i == 1 ? 4 : (i == 2 ? 7 : 9))) + i;
            }
            G[(i * 8 + r8) * JPITCH + j * 8 + c8] = acc;
        }
        if (cross)  // diagonal blocks from the cache
            for (int e = tid; e < 2 * JB * JB; e += ET) {
                const int h = e >> 8, r = (e >> 4) & 15, c = e & 15;
                G[(h * JB + r) * JPITCH + h * JB + c] = __ldcg((h ? DJ : DI) + (e & 255));
            }
    }
    __syncthreads();
    for (int e = tid; e < JP * JP; e += ET) {  // mirror the strictly-lower block triangle, W = I
        const int r = e / JP, c = e % JP;
        if (cross ? (r >= JB && c < JB) : ((r >> 3) > (c >> 3))) { const double2 v = G[c * JPITCH + r]; G[r * JPITCH + c] = make_double2(v.x, -v.y); }
        if (r == c) W[r * JPITCH + c] = make_double2(1.0, 0.0);
    }
    __syncthreads();
    auto mulph = [](double2 ph, double2 v) { return make_double2(ph.x * v.x - ph.y * v.y, ph.x * v.y + ph.y * v.x); };
    auto mulphc = [](double2 ph, double2 v) { return make_double2(ph.x * v.x + ph.y * v.y, ph.x * v.y - ph.y * v.x); };
    for (int sweep = 0; sweep < inner_sweeps; ++sweep) {
        if (tid == 0) s_sweep_any = 0;
        __syncthreads();
        // xrot (cached-diagonal visits of single-matrix launches): only the 16 x 16 pairs that join a column of block I
        // with a column of block J -- 16 bipartite steps instead of the 31 of the full tournament.  Pairs inside a block
        // are met in every one of the block's 63 visits of a sweep anyway; they keep being rotated in round 0 (full Gram)
        // and in the polishing sweeps.  The eigensolve is the critical path of a single-matrix round (39 of 68 us).
        const int nsteps = xrot ? JB : JP - 1;
        for (int step = 0; step < nsteps; ++step) {
            if (tid < JB) {
                int p, q;
                if (xrot) { p = tid; q = JB + ((tid + step) & (JB - 1)); }
                else if (tid == 0) { p = JP - 1; q = step; }
                else { p = (step + tid) % (JP - 1); q = (step - tid + (JP - 1)) % (JP - 1); }
                if (p > q) { int tt = p; p = q; q = tt; }
                const double alpha = G[p * JPITCH + p].x, beta = G[q * JPITCH + q].x;
                const double2 gg = G[p * JPITCH + q];
                const double ag2 = gg.x * gg.x + gg.y * gg.y;
                double c = 1.0, s = 0.0;
                double2 ph = make_double2(1.0, 0.0);
                const bool real_cols = s_cols[p] >= 0 && s_cols[q] >= 0;  // never touch padding slots
                if (!real_cols) {
                    // identity
                } else if (ag2 > tol * tol * fabs(alpha) * fabs(beta) && ag2 > 0.0 && alpha > dthr && beta > dthr) {
                    // inner rotation (|theta| <= pi/4) from two rsqrt levels: with a = (beta - alpha)/2, r = sqrt(a^2 + |g|^2):
                    // cos^2 = (1 + |a|/r)/2, sin = sign(a) |g| / (2 r cos); the phase needs a third, independent rsqrt
                    const double a = 0.5 * (beta - alpha);
                    const double rg = rsqrt(ag2), rr = rsqrt(a * a + ag2);
                    ph = make_double2(gg.x * rg, gg.y * rg);
                    const double c2 = 0.5 + 0.5 * fabs(a) * rr;
                    const double rc = rsqrt(c2);
                    c = c2 * rc;
                    s = 0.5 * (ag2 * rg) * rr * rc;
                    if (a < 0) s = -s;
                    // de Rijk ordering: keep the larger diagonal entry at the lower index (t = s / c)
                    const double t_ag = (s * rc) * (ag2 * rg);
                    if (alpha - t_ag < beta + t_ag) { const double cc = s, ss = -c; c = cc; s = ss; }
                    s_sweep_any = 1;
                    s_rot = 1;
                } else if (alpha < beta && beta > dthr) {
                    c = 0.0; s = -1.0;  // pure swap: a significant column moves in front of a smaller one
                    s_sweep_any = 1;
                }
                rot[tid] = make_double2(c, s);
                rph[tid] = ph;
                s_pq[2 * tid] = p;
                s_pq[2 * tid + 1] = q;
            }
            __syncthreads();
            {
                // fused two-sided update: a thread owns the 2x2 blocks G[{p1,q1}][{p2,q2}] of pairs (k1, k2) and
                // (k1 + 8, k2): G' = J1^H G J2 acting on rows (J1) and columns (J2)
                const int k2 = tid & 15;
                const double c2 = rot[k2].x, s2 = rot[k2].y;
                const bool id2 = (c2 == 1.0 && s2 == 0.0);
                const double2 ph2 = rph[k2];
                const int p2 = s_pq[2 * k2], q2 = s_pq[2 * k2 + 1];
                // ET = 128: two blocks per thread (k1, k1 + 8); 256: one block
                constexpr int NH = ET == 128 ? 2 : 1;
#pragma unroll
                for (int hh = 0; hh < NH; ++hh) {
                    const int k1 = (tid >> 4) + 8 * hh;
                    const double c1 = rot[k1].x, s1 = rot[k1].y;
                    const bool id1 = (c1 == 1.0 && s1 == 0.0);
                    if (id1 && id2) continue;
                    const double2 ph1 = rph[k1];
                    const int p1 = s_pq[2 * k1], q1 = s_pq[2 * k1 + 1];
                    double2 g00 = G[p1 * JPITCH + p2], g01 = G[p1 * JPITCH + q2];
                    double2 g10 = G[q1 * JPITCH + p2], g11 = G[q1 * JPITCH + q2];
                    if (!id1) {  // rows: y_p' = c y_p - s e^{i phi} y_q ; y_q' = s e^{-i phi} y_p + c y_q
                        const double2 e10 = mulph(ph1, g10), e11 = mulph(ph1, g11), f00 = mulphc(ph1, g00), f01 = mulphc(ph1, g01);
                        const double2 n00 = make_double2(c1 * g00.x - s1 * e10.x, c1 * g00.y - s1 * e10.y);
                        const double2 n01 = make_double2(c1 * g01.x - s1 * e11.x, c1 * g01.y - s1 * e11.y);
                        const double2 n10 = make_double2(s1 * f00.x + c1 * g10.x, s1 * f00.y + c1 * g10.y);
                        const double2 n11 = make_double2(s1 * f01.x + c1 * g11.x, s1 * f01.y + c1 * g11.y);
                        g00 = n00; g01 = n01; g10 = n10; g11 = n11;
                    }
                    if (!id2) {  // columns: x_p' = c x_p - s e^{-i phi} x_q ; x_q' = s e^{i phi} x_p + c x_q
                        const double2 e01 = mulphc(ph2, g01), e11 = mulphc(ph2, g11), f00 = mulph(ph2, g00), f10 = mulph(ph2, g10);
                        const double2 n00 = make_double2(c2 * g00.x - s2 * e01.x, c2 * g00.y - s2 * e01.y);
                        const double2 n10 = make_double2(c2 * g10.x - s2 * e11.x, c2 * g10.y - s2 * e11.y);
                        const double2 n01 = make_double2(s2 * f00.x + c2 * g01.x, s2 * f00.y + c2 * g01.y);
                        const double2 n11 = make_double2(s2 * f10.x + c2 * g11.x, s2 * f10.y + c2 * g11.y);
                        g00 = n00; g01 = n01; g10 = n10; g11 = n11;
                    }
                    G[p1 * JPITCH + p2] = g00; G[p1 * JPITCH + q2] = g01;
                    G[q1 * JPITCH + p2] = g10; G[q1 * JPITCH + q2] = g11;
                }
                if (!id2) {  // W <- W J (columns): rows (tid >> 4) + (ET / 16) h for pair k2
#pragma unroll
                    for (int h = 0; h < 512 / ET; ++h) {
                        const int r = (tid >> 4) + (ET / 16) * h;
                        const double2 xp = W[r * JPITCH + p2], xq = W[r * JPITCH + q2];
                        const double2 eq = mulphc(ph2, xq), ep = mulph(ph2, xp);
                        W[r * JPITCH + p2] = make_double2(c2 * xp.x - s2 * eq.x, c2 * xp.y - s2 * eq.y);
                        W[r * JPITCH + q2] = make_double2(s2 * ep.x + c2 * xq.x, s2 * ep.y + c2 * xq.y);
                    }
                }
            }
            __syncthreads();
        }
        const int any_now = s_sweep_any;
        if (any_now && tid == 0) s_any = 1;
        __syncthreads();
        if (!any_now) break;
    }
    const int any = s_any;
    if (tid == 0) {
        if (stat) atomicAdd(&stat[any ? 1 : 0], 1);
        if (any) { pr.last_mod[bi] = stamp; pr.last_mod[bj] = stamp; }
        else pr.last_ok[bi * nb + bj] = stamp;
        *flag = any;
        if (any && s_rot) rotated[b] = 1;  // pure re-ordering swaps do not keep the sweeps going
        if (any) atomicAdd(&nactive[b], 1);
    }
    // the block Grams after this visit: W^H G W restricted to each block (a full-Gram visit refreshes the cache even
    // when nothing was rotated)
    if (any || !cross)
        for (int e = tid; e < 2 * JB * JB; e += ET) {
            const int h = e >> 8, r = (e >> 4) & 15, c = e & 15;
            (h ? DJ : DI)[e & 255] = G[(h * JB + r) * JPITCH + h * JB + c];
        }
    if (!any) { trace_end(2, trace_t0); return; }
    double2* wout = Wbuf + (size_t)((size_t)b * maxpairs + pair) * (JP * JP);
    for (int e = tid; e < JP * JP; e += ET) wout[e] = W[(e >> 5) * JPITCH + (e & 31)];
    pdl_trigger();
    trace_end(2, trace_t0);
}

__global__ void __launch_bounds__(JTHREADS, 512 / JTHREADS)
jacobi_update_kernel(const SvdProblem* __restrict__ probs, int round, const int* __restrict__ rotated, int SU, int maxpairs,
                     const double2* __restrict__ Wbuf, const int* __restrict__ pflag) {
    pdl_wait();
    const int b = blockIdx.y, pair = blockIdx.x / SU, chunk = blockIdx.x - pair * SU;
    if (rotated[b] < 0 || pflag[(size_t)b * maxpairs + pair] == 0) return;
    const SvdProblem pr = probs[b];
    int bi, bj;
    if (!round_pair(pr, round, pair, bi, bj)) return;
    const unsigned long long trace_t0 = d_trace_buf ? trace_now() : 0ull;
    __shared__ double2 Ws[JP * JPITCH];
    constexpr int SP = JP + 4;                 // pitch of the sum plane: conflict-free LDS.64 for the B fragments
    __shared__ double Wsum[JP * SP];           // re + im of W: the third operand of the 3-multiplication product
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    {
        const double2* wsrc = Wbuf + (size_t)((size_t)b * maxpairs + pair) * (JP * JP);
        for (int e = tid; e < JP * JP; e += JTHREADS) {
            const double2 wv = __ldcg(wsrc + e);
            Ws[(e >> 5) * JPITCH + (e & 31)] = wv;
            Wsum[(e >> 5) * SP + (e & 31)] = wv.x + wv.y;
        }
    }
    int colk[8], colo[4][2];  // global columns of this lane's A-fragment slots (k = 4 k4 + t) and output slots
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) colk[k4] = panel_col(k4 * 4 + t, bi, bj, pr.n);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) colo[j][q] = panel_col(j * 8 + 2 * t + q, bi, bj, pr.n);
    __syncthreads();
    // row groups of 8: first those of A (m rows), then those of V (n rows) when V is accumulated
    const int ga = (pr.m + 7) >> 3, gv = pr.V ? (pr.n + 7) >> 3 : 0, gt = ga + gv;
    const int gper = (gt + SU - 1) / SU, c0 = chunk * gper, c1 = min(gt, c0 + gper);
    for (int gi = c0 + warp; gi < c1; gi += JTHREADS / 32) {
        double2* base = gi < ga ? pr.A : pr.V;
        const int nrows = gi < ga ? pr.m : pr.n;
        const int row = (gi < ga ? gi : gi - ga) * 8 + g;
        const bool rok = row < nrows;
        double2 a[8];
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4)
            a[k4] = (rok && colk[k4] >= 0) ? __ldcg(base + (size_t)colk[k4] * nrows + row) : make_double2(0.0, 0.0);
        // 3-multiplication complex product: P1 = sum ar br, P2 = sum ai bi, P3 = sum (ar + ai)(br + bi);
        // re = P1 - P2, im = P3 - P1 - P2 (96 DMMAs per 8 x 32 tile instead of 128; the rotations stay unitary to
        // rounding either way, and the Jacobi iteration corrects what the different rounding leaves)
        double p1[4][2], p2[4][2], p3[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) p1[j][0] = p1[j][1] = p2[j][0] = p2[j][1] = p3[j][0] = p3[j][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const double asum = a[k4].x + a[k4].y;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 bb = Ws[(k4 * 4 + t) * JPITCH + j * 8 + g];
                const double bs = Wsum[(k4 * 4 + t) * SP + j * 8 + g];
                dmma(p1[j][0], p1[j][1], a[k4].x, bb.x);
                dmma(p2[j][0], p2[j][1], a[k4].y, bb.y);
                dmma(p3[j][0], p3[j][1], asum, bs);
            }
        }
        if (rok) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (colo[j][q] >= 0)
                        base[(size_t)colo[j][q] * nrows + row] = make_double2(p1[j][q] - p2[j][q], p3[j][q] - p1[j][q] - p2[j][q]);
        }
    }
    pdl_trigger();
    trace_end(3, trace_t0);
}

// =============================================================================================================
// Dataflow sweep kernel (batches with enough block pairs per round to fill the machine).
//
// The three-kernel round above synchronises ALL block pairs of a sub-batch at every phase boundary, 63 times per
// sweep: the latency-bound eigensolves leave the tensor pipe idle, every kernel has a partial last wave, and the
// phases of different sub-batches only overlap by luck.  But a block pair (I, J) of round r+1 depends on exactly two
// pairs of round r -- the ones that held I and J.  This kernel expresses that directly:
//   * the (matrix, round, pair) visits of a sweep form a task list sorted by round; persistent CTAs (4 per SM) draw
//     tasks from an atomic counter;
//   * a task waits until both its blocks carry version == round (ld.acquire spin by one thread), runs
//     Gram -> eigensolve -> update for its pair entirely inside the CTA (G and W never leave shared memory; no partial
//     Grams, no W buffer, no apply flags in global memory), then publishes version = round + 1 for both blocks
//     (__threadfence + st.release);
//   * CTAs of one SM are in different phases at any time, so the tensor pipe sees the Gram / update DMMAs of some
//     tasks while others sit in their eigensolve -- without any barrier wider than a CTA.
// Tasks are drawn in list order, so a task's producers (previous round) were drawn earlier by CTAs that are running:
// no deadlock, no co-residency requirement.  Everything a task reads that another CTA may have written (A, V, D,
// stamps, versions) is read with ld.cg / ld.acquire (L1 is not coherent across SMs).
// FULL = true: full 32 x 32 Gram (round 0 of a sweep, polishing sweeps); false: cross block only + cached diagonal blocks.
// =============================================================================================================
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ldcg_int(const int* p) { return __ldcg(p); }

constexpr int FLOW_THREADS = 128;
constexpr int FSP = JP + 4;  // pitch of the re + im plane of W (conflict-free LDS.64 for the B fragments)

// Shared-memory arena of one CTA (dynamic): [ W | Wsum ][ G ][ extra ].  During the Gram main loop the WHOLE arena is a
// ring of per-thread cp.async slots; the update phase uses [ G | extra ] the same way (G is dead after the eigensolve).
constexpr int FLOW_W_ELEMS = JP * JPITCH + JP * FSP / 2;   // W (double2) + Wsum (double) in double2 units
constexpr int FLOW_G_ELEMS = JP * JPITCH;
template <int MINB> struct FlowCfg {
    static constexpr int EXTRA = MINB >= 4 ? 0 : (MINB == 3 ? 1536 : 4096);                    // extra ring space, double2 units
    static constexpr int ARENA = FLOW_W_ELEMS + FLOW_G_ELEMS + EXTRA;
    static constexpr int DG = ARENA / (4 * FLOW_THREADS);                                      // Gram ring depth (4-row chunks)
    static constexpr int DU = (FLOW_G_ELEMS + EXTRA) / (8 * FLOW_THREADS);                     // update ring depth (8-row groups)
    static constexpr size_t SMEM = (size_t)ARENA * sizeof(double2);
};

template <bool FULL, int MINB>
__global__ void __launch_bounds__(FLOW_THREADS, MINB)
jacobi_flow_kernel(const SvdProblem* __restrict__ probs, const uint2* __restrict__ tasks, int ntasks, int* __restrict__ counter,
                   double tol, int* __restrict__ rotated, const double* __restrict__ fro2, int inner_sweeps, int* __restrict__ stat,
                   int stamp_base, int* __restrict__ nactive, int* __restrict__ errflag, int xrot) {
    typedef FlowCfg<MINB> Cfg;
    extern __shared__ __align__(16) unsigned char flow_smem[];
    double2* const arena = reinterpret_cast<double2*>(flow_smem);
    double2* const W = arena;
    double* const Wsum = reinterpret_cast<double*>(arena + JP * JPITCH);
    double2* const G = arena + FLOW_W_ELEMS;
    __shared__ double2 rot[JB], rph[JB];
    __shared__ int s_pq[2 * JB];
    __shared__ int s_any, s_sweep_any, s_rot, s_task;
    __shared__ int s_cols[JP];
    static_assert(sizeof(double2) * FLOW_W_ELEMS >= sizeof(double) * 2 * 40 * 32, "reduction scratch");
    static_assert(Cfg::DG >= 2 && Cfg::DU >= 1, "ring depths");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (;;) {
        __syncthreads();  // every thread is done with the previous task's shared state (incl. s_task)
        if (tid == 0) s_task = atomicAdd(counter, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= ntasks) return;
        const uint2 tk = __ldg(tasks + task);
        const int b = (int)tk.x, round = (int)(tk.y >> 16), pair = (int)(tk.y & 0xffffu);
        const SvdProblem* __restrict__ prp = probs + b;   // fields are re-read where needed (registers are the scarce resource here)
        const int nb = prp->nblocks, pn = prp->n;
        int bi, bj;  // every block is in exactly one pair per round (nb is even): versions advance even for padding / idle pairs
        if (pair == 0) { bi = nb - 1; bj = round; }
        else { bi = (round + pair) % (nb - 1); bj = (round - pair + (nb - 1)) % (nb - 1); }
        if (bi > bj) { const int tt = bi; bi = bj; bj = tt; }
        if (tid == 0) {
            unsigned ns = 32, polls = 0;
            while (ld_acquire(prp->ver + bi) < round || ld_acquire(prp->ver + bj) < round) {
                __nanosleep(ns);
                if (ns < 1024) ns *= 2;
                // watchdog (~1 s): a producer that never publishes is a bug, not a reason to hang the device; the host
                // reports QTN_ECUDA when the flag is set
                if (++polls > (1u << 20)) { atomicExch(errflag, 1); break; }
            }
        }
        __syncthreads();
        const bool live = __ldcg(rotated + b) >= 0 && bi * JB < pn;
        const bool idle = live && max(ldcg_int(prp->last_mod + bi), ldcg_int(prp->last_mod + bj)) < ldcg_int(prp->last_ok + bi * nb + bj);
        if (!live || idle) {
            if (tid == 0) {
                if (live && stat) atomicAdd(&stat[0], 1);
                st_release(prp->ver + bi, round + 1);
                st_release(prp->ver + bj, round + 1);
            }
            continue;
        }
        const int m = prp->m;
        const int stamp = stamp_base + round;
        // ------------------------------------------------ Gram of the panel -> G ------------------------------
        unsigned long long trace_t0 = d_trace_buf ? trace_now() : 0ull;
        {
            const double2* cp[4];
            bool cv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int col = panel_col(i * 8 + g, bi, bj, pn);
                cv[i] = col >= 0;
                cp[i] = prp->A + (size_t)(col >= 0 ? col : 0) * m;
            }
            const int ngroups = (m + 15) >> 4;
            constexpr int NWP = FLOW_THREADS / 32;
            double* red = reinterpret_cast<double*>(W);  // [2][NB4][32]
            if constexpr (FULL) {
                double gr[10][2], gi[10][2];
#pragma unroll
                for (int k = 0; k < 10; ++k) gr[k][0] = gr[k][1] = gi[k][0] = gi[k][1] = 0.0;
                for (int grp = warp; grp < ngroups; grp += NWP) {
                    double2 f[4][4];
                    const int r = grp * 16 + t;
#pragma unroll
                    for (int s = 0; s < 4; ++s)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int row = r + 4 * s;
                            f[s][i] = (cv[i] && row < m) ? __ldcg(cp[i] + row) : make_double2(0.0, 0.0);
                        }
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        int k = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = i; j < 4; ++j, ++k) {
                                dmma(gr[k][0], gr[k][1], f[s][i].x, f[s][j].x);
                                dmma(gr[k][0], gr[k][1], f[s][i].y, f[s][j].y);
                                dmma(gi[k][0], gi[k][1], f[s][i].x, f[s][j].y);
                                dmma(gi[k][0], gi[k][1], -f[s][i].y, f[s][j].x);
                            }
                    }
                }
#pragma unroll
                for (int stage = 0; stage < 2; ++stage) {  // fixed-order tree over the 4 warps (deterministic)
                    const int half = stage == 0 ? 2 : 1;
                    if (warp >= half && warp < 2 * half) {
#pragma unroll
                        for (int k = 0; k < 10; ++k) {
                            red[((warp - half) * 40 + 4 * k + 0) * 32 + lane] = gr[k][0];
                            red[((warp - half) * 40 + 4 * k + 1) * 32 + lane] = gr[k][1];
                            red[((warp - half) * 40 + 4 * k + 2) * 32 + lane] = gi[k][0];
                            red[((warp - half) * 40 + 4 * k + 3) * 32 + lane] = gi[k][1];
                        }
                    }
                    __syncthreads();
                    if (warp < half) {
#pragma unroll
                        for (int k = 0; k < 10; ++k) {
                            gr[k][0] += red[(warp * 40 + 4 * k + 0) * 32 + lane];
                            gr[k][1] += red[(warp * 40 + 4 * k + 1) * 32 + lane];
                            gi[k][0] += red[(warp * 40 + 4 * k + 2) * 32 + lane];
                            gi[k][1] += red[(warp * 40 + 4 * k + 3) * 32 + lane];
                        }
                    }
                    __syncthreads();
                }
                if (warp == 0) {  // upper block triangle: block k = (i, j >= i), lane (g, t) owns row g, columns 2t, 2t+1
                    int k = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = i; j < 4; ++j, ++k) {
                            G[(i * 8 + g) * JPITCH + j * 8 + 2 * t] = make_double2(gr[k][0], gi[k][0]);
                            G[(i * 8 + g) * JPITCH + j * 8 + 2 * t + 1] = make_double2(gr[k][1], gi[k][1]);
                        }
                }
            } else {
                double m1[4][2], m2[4][2], m3[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) m1[k][0] = m1[k][1] = m2[k][0] = m2[k][1] = m3[k][0] = m3[k][1] = 0.0;
                // 4-row chunks (one k-step of the DMMA) through a DG-deep ring of per-thread cp.async slots spanning the whole
                // arena: a thread reads back exactly the four fragments it fetched itself, so the ring needs no barrier.
                // (ncu: with one chunk of register prefetch this phase was 21 % of the kernel's time, all long-scoreboard,
                // for 6 us of tensor-pipe work: only ~2 CTAs of an SM are in a tensor phase at any time.)
                const int nch = (m + 3) >> 2;
                const int nk = nch > warp ? (nch - warp + NWP - 1) / NWP : 0;   // this warp's chunks: warp + NWP * k
                double2* const gslot = arena + tid;                             // [slot][i][FLOW_THREADS]
                auto fetch = [&](int kq) {
                    if (kq < nk) {
                        const int row = (warp + NWP * kq) * 4 + t;
                        double2* dst = gslot + (size_t)(kq % Cfg::DG) * (4 * FLOW_THREADS);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool ok = cv[i] && row < m;
                            cp_async16(dst + i * FLOW_THREADS, ok ? (const void*)(cp[i] + row) : (const void*)cp[i], ok);
                        }
                    }
                    cp_async_commit();
                };
                for (int kq = 0; kq < Cfg::DG - 1; ++kq) fetch(kq);
                for (int kq = 0; kq < nk; ++kq) {
                    fetch(kq + Cfg::DG - 1);          // overwrites the slot read in iteration kq - 1
                    cp_async_wait<Cfg::DG - 1>();     // group kq has landed
                    const double2* src = gslot + (size_t)(kq % Cfg::DG) * (4 * FLOW_THREADS);
                    double2 f[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[i] = src[i * FLOW_THREADS];
                    asm volatile("" ::: "memory");
                    const double sa0 = f[0].x + f[0].y, sa1 = f[1].x + f[1].y;
                    const double db2 = f[2].x - f[2].y, db3 = f[3].x - f[3].y;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int kk = 2 * i + j;
                            dmma(m1[kk][0], m1[kk][1], f[i].x, f[2 + j].x);
                            dmma(m2[kk][0], m2[kk][1], f[i].y, f[2 + j].y);
                            dmma(m3[kk][0], m3[kk][1], i ? sa1 : sa0, j ? db3 : db2);
                        }
                }
                cp_async_wait<0>();
                __syncthreads();   // every warp is done with its ring slots: the reduction scratch and G overlay them
                double gr[4][2], gi[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int q = 0; q < 2; ++q) { gr[k][q] = m1[k][q] + m2[k][q]; gi[k][q] = m1[k][q] - m2[k][q] - m3[k][q]; }
#pragma unroll
                for (int stage = 0; stage < 2; ++stage) {
                    const int half = stage == 0 ? 2 : 1;
                    if (warp >= half && warp < 2 * half) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            red[((warp - half) * 16 + 4 * k + 0) * 32 + lane] = gr[k][0];
                            red[((warp - half) * 16 + 4 * k + 1) * 32 + lane] = gr[k][1];
                            red[((warp - half) * 16 + 4 * k + 2) * 32 + lane] = gi[k][0];
                            red[((warp - half) * 16 + 4 * k + 3) * 32 + lane] = gi[k][1];
                        }
                    }
                    __syncthreads();
                    if (warp < half) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            gr[k][0] += red[(warp * 16 + 4 * k + 0) * 32 + lane];
                            gr[k][1] += red[(warp * 16 + 4 * k + 1) * 32 + lane];
                            gi[k][0] += red[(warp * 16 + 4 * k + 2) * 32 + lane];
                            gi[k][1] += red[(warp * 16 + 4 * k + 3) * 32 + lane];
                        }
                    }
                    __syncthreads();
                }
                if (warp == 0) {  // block k = 2 i + j' is block (i, 2 + j') of the panel Gram
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = k >> 1, j = 2 + (k & 1);
                        G[(i * 8 + g) * JPITCH + j * 8 + 2 * t] = make_double2(gr[k][0], gi[k][0]);
                        G[(i * 8 + g) * JPITCH + j * 8 + 2 * t + 1] = make_double2(gr[k][1], gi[k][1]);
                    }
                }
            }
        }
        trace_end(FULL ? 0 : 1, trace_t0);
        trace_t0 = d_trace_buf ? trace_now() : 0ull;
        // ------------------------------------------------ eigensolve of G, rotations accumulated in W ---------
        const double dthr = 1e-30 * fro2[b];
        double2* DI = prp->D + (size_t)bi * (JB * JB);
        double2* DJ = prp->D + (size_t)bj * (JB * JB);
        if (tid < JP) s_cols[tid] = panel_col(tid, bi, bj, pn);
        if (tid == 0) { s_any = 0; s_rot = 0; }
        if (!FULL)  // diagonal blocks from the cache
            for (int e = tid; e < 2 * JB * JB; e += FLOW_THREADS) {
                const int h = e >> 8, r = (e >> 4) & 15, c = e & 15;
                G[(h * JB + r) * JPITCH + h * JB + c] = __ldcg((h ? DJ : DI) + (e & 255));
            }
        __syncthreads();  // (also: the reduction scratch in W / Wsum is dead from here on)
        for (int e = tid; e < JP * JP; e += FLOW_THREADS) {  // mirror the strictly-lower block triangle, W = I
            const int r = e / JP, c = e % JP;
            if (!FULL ? (r >= JB && c < JB) : ((r >> 3) > (c >> 3))) { const double2 v = G[c * JPITCH + r]; G[r * JPITCH + c] = make_double2(v.x, -v.y); }
            W[r * JPITCH + c] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
        for (int sweep = 0; sweep < inner_sweeps; ++sweep) {
            if (tid == 0) s_sweep_any = 0;
            __syncthreads();
            const int nsteps = (!FULL && xrot) ? JB : JP - 1;   // cached-diagonal visits: cross pairs only (see jacobi_eig_kernel)
            for (int step = 0; step < nsteps; ++step) {
                if (tid < JB) {
                    int p, q;
                    if (!FULL && xrot) { p = tid; q = JB + ((tid + step) & (JB - 1)); }
                    else if (tid == 0) { p = JP - 1; q = step; }
                    else { p = (step + tid) % (JP - 1); q = (step - tid + (JP - 1)) % (JP - 1); }
                    if (p > q) { int tt = p; p = q; q = tt; }
                    const double alpha = G[p * JPITCH + p].x, beta = G[q * JPITCH + q].x;
                    const double2 gg = G[p * JPITCH + q];
                    const double ag2 = gg.x * gg.x + gg.y * gg.y;
                    double c = 1.0, s = 0.0;
                    double2 ph = make_double2(1.0, 0.0);
                    const bool real_cols = s_cols[p] >= 0 && s_cols[q] >= 0;  // never touch padding slots
                    if (!real_cols) {
                        // identity
                    } else if (ag2 > tol * tol * fabs(alpha) * fabs(beta) && ag2 > 0.0 && alpha > dthr && beta > dthr) {
                        const double a = 0.5 * (beta - alpha);
                        const double rg = rsqrt(ag2), rr = rsqrt(a * a + ag2);
                        ph = make_double2(gg.x * rg, gg.y * rg);
                        const double c2 = 0.5 + 0.5 * fabs(a) * rr;
                        const double rc = rsqrt(c2);
                        c = c2 * rc;
                        s = 0.5 * (ag2 * rg) * rr * rc;
                        if (a < 0) s = -s;
                        const double t_ag = (s * rc) * (ag2 * rg);
                        if (alpha - t_ag < beta + t_ag) { const double cc = s, ss = -c; c = cc; s = ss; }
                        s_sweep_any = 1;
                        s_rot = 1;
                    } else if (alpha < beta && beta > dthr) {
                        c = 0.0; s = -1.0;
                        s_sweep_any = 1;
                    }
                    rot[tid] = make_double2(c, s);
                    rph[tid] = make_double2(s * ph.x, s * ph.y);   // S = s e^{i phi}: the phase is folded into the sine once
                    s_pq[2 * tid] = p;
                    s_pq[2 * tid + 1] = q;
                }
                __syncthreads();
                {
                    // u' = c u - S v  and  v' = conj(S) u + c v  (rows; columns take conj(S) / S): 6 FP64 instructions per
                    // complex output instead of 8 -- the eigensolve's DFMAs queue behind the DMMAs of the other CTAs
                    auto lin_m = [](double c, double2 u, double2 S, double2 v) {   // c u - S v
                        return make_double2(fma(-S.x, v.x, fma(S.y, v.y, c * u.x)), fma(-S.x, v.y, fma(-S.y, v.x, c * u.y)));
                    };
                    auto lin_mc = [](double c, double2 u, double2 S, double2 v) {  // c u - conj(S) v
                        return make_double2(fma(-S.x, v.x, fma(-S.y, v.y, c * u.x)), fma(-S.x, v.y, fma(S.y, v.x, c * u.y)));
                    };
                    auto lin_p = [](double c, double2 u, double2 S, double2 v) {   // c u + S v
                        return make_double2(fma(S.x, v.x, fma(-S.y, v.y, c * u.x)), fma(S.x, v.y, fma(S.y, v.x, c * u.y)));
                    };
                    auto lin_pc = [](double c, double2 u, double2 S, double2 v) {  // c u + conj(S) v
                        return make_double2(fma(S.x, v.x, fma(S.y, v.y, c * u.x)), fma(S.x, v.y, fma(-S.y, v.x, c * u.y)));
                    };
                    const int k2 = tid & 15;
                    const double c2 = rot[k2].x;
                    const double2 S2 = rph[k2];
                    const bool id2 = (c2 == 1.0 && rot[k2].y == 0.0);
                    const int p2 = s_pq[2 * k2], q2 = s_pq[2 * k2 + 1];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int k1 = (tid >> 4) + 8 * hh;
                        const double c1 = rot[k1].x;
                        const bool id1 = (c1 == 1.0 && rot[k1].y == 0.0);
                        if (id1 && id2) continue;
                        const double2 S1 = rph[k1];
                        const int p1 = s_pq[2 * k1], q1 = s_pq[2 * k1 + 1];
                        double2 g00 = G[p1 * JPITCH + p2], g01 = G[p1 * JPITCH + q2];
                        double2 g10 = G[q1 * JPITCH + p2], g11 = G[q1 * JPITCH + q2];
                        if (!id1) {  // rows: y_p' = c y_p - S y_q ; y_q' = conj(S) y_p + c y_q
                            const double2 n00 = lin_m(c1, g00, S1, g10), n01 = lin_m(c1, g01, S1, g11);
                            const double2 n10 = lin_pc(c1, g10, S1, g00), n11 = lin_pc(c1, g11, S1, g01);
                            g00 = n00; g01 = n01; g10 = n10; g11 = n11;
                        }
                        if (!id2) {  // columns: x_p' = c x_p - conj(S) x_q ; x_q' = S x_p + c x_q
                            const double2 n00 = lin_mc(c2, g00, S2, g01), n10 = lin_mc(c2, g10, S2, g11);
                            const double2 n01 = lin_p(c2, g01, S2, g00), n11 = lin_p(c2, g11, S2, g10);
                            g00 = n00; g01 = n01; g10 = n10; g11 = n11;
                        }
                        G[p1 * JPITCH + p2] = g00; G[p1 * JPITCH + q2] = g01;
                        G[q1 * JPITCH + p2] = g10; G[q1 * JPITCH + q2] = g11;
                    }
                    if (!id2) {  // W <- W J (columns), rows (tid >> 4) + 8 h for pair k2
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const int r = (tid >> 4) + 8 * h;
                            const double2 xp = W[r * JPITCH + p2], xq = W[r * JPITCH + q2];
                            W[r * JPITCH + p2] = lin_mc(c2, xp, S2, xq);
                            W[r * JPITCH + q2] = lin_p(c2, xq, S2, xp);
                        }
                    }
                }
                __syncthreads();
            }
            const int any_now = s_sweep_any;
            if (any_now && tid == 0) s_any = 1;
            __syncthreads();
            if (!any_now) break;
        }
        const int any = s_any;
        if (tid == 0) {
            if (stat) atomicAdd(&stat[any ? 1 : 0], 1);
            if (any) { prp->last_mod[bi] = stamp; prp->last_mod[bj] = stamp; }
            else prp->last_ok[bi * nb + bj] = stamp;
            if (any && s_rot) rotated[b] = 1;
            if (any) atomicAdd(&nactive[b], 1);
        }
        if (any || FULL)
            for (int e = tid; e < 2 * JB * JB; e += FLOW_THREADS) {
                const int h = e >> 8, r = (e >> 4) & 15, c = e & 15;
                (h ? DJ : DI)[e & 255] = G[(h * JB + r) * JPITCH + h * JB + c];
            }
        trace_end(2, trace_t0);
        // ------------------------------------------------ update  P <- P W  (A panel, then V panel) -----------
        if (any) {
            trace_t0 = d_trace_buf ? trace_now() : 0ull;
            for (int e = tid; e < JP * JP; e += FLOW_THREADS) {
                const double2 wv = W[(e >> 5) * JPITCH + (e & 31)];
                Wsum[(e >> 5) * FSP + (e & 31)] = wv.x + wv.y;
            }
            __syncthreads();   // Wsum is complete; every thread has read its D write-back values out of G
            // Software pipeline: the 8 fragments of the next DU row groups are fetched by cp.async into per-thread slots of
            // [ G | extra ] (G is dead after the eigensolve) while the 96 DMMAs of the current group run; every thread reads
            // back exactly what it fetched itself.  The W fragments of k4 + 1 are loaded under the DMMAs of k4.  All column
            // addressing is hoisted out of the row loop (a warp holds one DMMA issue slot in four: every other instruction
            // of the loop is time the tensor pipe is not fed -- tools/micro/dmma_warp.cu).
            double2* const slot = G + tid;   // [slot][k4][FLOW_THREADS]
            constexpr int NWP = FLOW_THREADS / 32;
            constexpr int NJ = MINB >= 4 ? 2 : 4;   // column blocks per pass (3 NJ accumulator pairs)
            auto run_panel = [&](double2* __restrict__ base, const int nrows) {
                const int ngr = (nrows + 7) >> 3;
                const int ngw = ngr > warp ? (ngr - warp + NWP - 1) / NWP : 0;   // this warp's row groups: warp + NWP * k
                const double2* src_col[8];   // fragment slot k4 of this lane: column k4 * 4 + t, row g of the group
                unsigned src_ok = 0;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const int col = panel_col(k4 * 4 + t, bi, bj, pn);
                    src_col[k4] = base + (size_t)(col >= 0 ? col : 0) * nrows + g;
                    src_ok |= (col >= 0 ? 1u : 0u) << k4;
                }
                auto fetch_rows = [&](int kq) {
                    if (kq < ngw) {
                        const int r0 = (warp + NWP * kq) * 8;
                        const bool rok = r0 + g < nrows;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(slot + (size_t)(kq % Cfg::DU) * (8 * FLOW_THREADS));
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4) {
                            const int sz = (rok && ((src_ok >> k4) & 1u)) ? 16 : 0;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + (unsigned)(k4 * FLOW_THREADS * 16)),
                                         "l"(sz ? src_col[k4] + r0 : src_col[k4]), "r"(sz));
                        }
                    }
                    cp_async_commit();
                };
                for (int kq = 0; kq < Cfg::DU; ++kq) fetch_rows(kq);
                for (int kq = 0; kq < ngw; ++kq) {
                    double2 a[8];
                    cp_async_wait<Cfg::DU - 1>();   // group kq has landed
                    const double2* src = slot + (size_t)(kq % Cfg::DU) * (8 * FLOW_THREADS);
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) a[k4] = src[k4 * FLOW_THREADS];
                    asm volatile("" ::: "memory");   // the slot is re-filled only after it has been read
                    fetch_rows(kq + Cfg::DU);
                    const int row = (warp + NWP * kq) * 8 + g;
                    const bool rok = row < nrows;
#pragma unroll 1
                    for (int jh = 0; jh < 4 / NJ; ++jh) {
                        double p1[NJ][2], p2[NJ][2], p3[NJ][2];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) p1[j][0] = p1[j][1] = p2[j][0] = p2[j][1] = p3[j][0] = p3[j][1] = 0.0;
                        const double2* wp = W + t * JPITCH + (NJ * jh) * 8 + g;
                        const double* sp = Wsum + t * FSP + (NJ * jh) * 8 + g;
                        double2 b0[NJ], b1[NJ];
                        double s0[NJ], s1[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) { b0[j] = wp[j * 8]; s0[j] = sp[j * 8]; }
#pragma unroll
                        for (int k4 = 0; k4 < 8; k4 += 2) {
#pragma unroll
                            for (int j = 0; j < NJ; ++j) { b1[j] = wp[(k4 + 1) * 4 * JPITCH + j * 8]; s1[j] = sp[(k4 + 1) * 4 * FSP + j * 8]; }
                            {
                                const double asum = a[k4].x + a[k4].y;
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma(p1[j][0], p1[j][1], a[k4].x, b0[j].x);
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma(p2[j][0], p2[j][1], a[k4].y, b0[j].y);
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma(p3[j][0], p3[j][1], asum, s0[j]);
                            }
                            if (k4 + 2 < 8) {
#pragma unroll
                                for (int j = 0; j < NJ; ++j) { b0[j] = wp[(k4 + 2) * 4 * JPITCH + j * 8]; s0[j] = sp[(k4 + 2) * 4 * FSP + j * 8]; }
                            }
                            {
                                const double asum = a[k4 + 1].x + a[k4 + 1].y;
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma(p1[j][0], p1[j][1], a[k4 + 1].x, b1[j].x);
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma(p2[j][0], p2[j][1], a[k4 + 1].y, b1[j].y);
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma(p3[j][0], p3[j][1], asum, s1[j]);
                            }
                        }
                        if (rok) {
                            // output slot (j, q) of this lane: column (NJ jh + j) * 8 + 2 t + q.  Columns 2t, 2t+1 are neighbours
                            // of one block (JB is even), so one panel_col per j locates both.
#pragma unroll
                            for (int j = 0; j < NJ; ++j) {
                                const int col = panel_col((NJ * jh + j) * 8 + 2 * t, bi, bj, pn);
                                if (col >= 0) {
                                    double2* o = base + (size_t)col * nrows + row;
                                    o[0] = make_double2(p1[j][0] - p2[j][0], p3[j][0] - p1[j][0] - p2[j][0]);
                                    if (col + 1 < pn) o[nrows] = make_double2(p1[j][1] - p2[j][1], p3[j][1] - p1[j][1] - p2[j][1]);
                                }
                            }
                        }
                    }
                }
                cp_async_wait<0>();
            };
            run_panel(prp->A, m);
            if (prp->V) run_panel(prp->V, pn);
            trace_end(3, trace_t0);
        }
        // publish: every thread's stores (A, V, D, stamps) are fenced before the versions move
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            st_release(prp->ver + bi, round + 1);
            st_release(prp->ver + bj, round + 1);
        }
    }
}

__global__ void set_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// fro2[b] = ||A_b||_F^2
__global__ void fro_norm_kernel(const SvdProblem* probs, double* fro2) {
    const SvdProblem pr = probs[blockIdx.y];
    const size_t tot = (size_t)pr.m * pr.n;
    double s = 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const double2 v = pr.A[e];
        s += v.x * v.x + v.y * v.y;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&fro2[blockIdx.y], s);
}

__global__ void set_identity_kernel(const SvdProblem* probs) {
    const SvdProblem pr = probs[blockIdx.y];
    if (!pr.V) return;
    const size_t nn = (size_t)pr.n * pr.n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x)
        pr.V[e] = make_double2((e % pr.n) == (e / pr.n) ? 1.0 : 0.0, 0.0);
}

// Columns with ||a_j||^2 <= kDeadFactor * ||A||_F^2 were deflated by the rotation kernel (threshold 1e-30, with
// slack for the recomputed norm).
constexpr double kDeadFactor = 4e-30;

// sigma_j = ||a_j||  (one warp per column)
__global__ void column_norms_kernel(const SvdProblem* probs, double* const* sig) {
    const SvdProblem pr = probs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    for (int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); col < pr.n; col += gridDim.x * (blockDim.x >> 5)) {
        const double2* a = pr.A + (size_t)col * pr.m;
        // scaled accumulation is unnecessary here: |a| <= ||A||_F which the callers keep O(1..1e150)
        double s = 0;
        for (int r = lane; r < pr.m; r += 32) s += a[r].x * a[r].x + a[r].y * a[r].y;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sig[blockIdx.y][col] = sqrt(s);
    }
}

struct FinishArgs {
    double2* U;    // m x r
    double* S;     // r
    double2* Vh;   // r x n
    int64_t* k;    // kept count
    double* disc;  // 2-norm of the discarded tail (may be null)
    int transposed;  // the Jacobi ran on A^H: U <-> V roles swap
    int m0, n0;      // caller's shape
};

// rank[j] = position of sigma_j in descending order (stable); S sorted; truncation rule (K6).
// dead[b] = number of columns at or below the deflation threshold (never orthogonalised: their direction is
// rounding noise); they sort last and are replaced by an orthonormal completion afterwards.
__global__ void sort_truncate_kernel(const SvdProblem* probs, double* const* sig, int* const* rank,
                                     const FinishArgs* fin, double er, long long maxdim, const double* __restrict__ fro2,
                                     int* __restrict__ dead) {
    const SvdProblem pr = probs[blockIdx.x];
    const FinishArgs f = fin[blockIdx.x];
    const double* s = sig[blockIdx.x];
    int* rk = rank[blockIdx.x];
    const int n = pr.n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double sj = s[j];
        int r = 0;
        for (int k = 0; k < n; ++k) r += (s[k] > sj) || (s[k] == sj && k < j);
        rk[j] = r;
        f.S[r] = sj;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // src/svd.jl:29-33: cumsum of the reversed squares, first tail > er (strict)
        long long k = 0;
        if (er < 0) k = n;
        else {
            double acc = 0;
            for (int r = 1; r <= n; ++r) {
                const double v = f.S[n - r];
                acc += v * v;
                if (sqrt(acc) > er) { k = n - r + 1; break; }
            }
        }
        if (maxdim > 0 && k > maxdim) k = maxdim;
        double d = 0;
        for (int j = n - 1; j >= (int)k; --j) d += f.S[j] * f.S[j];
        *f.k = k;
        if (f.disc) *f.disc = sqrt(d);
        int nd = 0;
        const double thr = kDeadFactor * fro2[blockIdx.x];
        for (int j = n - 1; j >= 0 && f.S[j] * f.S[j] <= thr; --j) ++nd;
        dead[blockIdx.x] = nd;
    }
}

// U[:, rank[j]] = a_j / sigma_j ; Vh[rank[j], :] = conj(V[:, j])   (roles swapped if transposed).
// A deflated column (norm below the rotation threshold, never orthogonalised against the others: pure rounding
// noise of a rank-deficient A) must not survive as a unit vector -- callers that form S*Vh = U^H A would pick up
// a direction that is not orthogonal to the rest.  It is written as zero here and replaced by an orthonormal
// completion (complete_null_vectors) afterwards, so the factor is an isometry like LAPACK's.
__global__ void scatter_factors_kernel(const SvdProblem* probs, double* const* sig, int* const* rank, const FinishArgs* fin,
                                       const double* __restrict__ fro2) {
    const SvdProblem pr = probs[blockIdx.y];
    const FinishArgs f = fin[blockIdx.y];
    const int n = pr.n, m = pr.m;
    const int* rk = rank[blockIdx.y];
    const double* s = sig[blockIdx.y];
    const size_t tot = (size_t)(m + n) * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(e / (m + n)), r = (int)(e % (m + n)), dst = rk[j];
        if (r < m) {
            const double sj = s[j];
            double2 v = pr.A[(size_t)j * m + r];
            const bool dead = sj * sj <= kDeadFactor * fro2[blockIdx.y];
            if (sj > 0 && !dead) { v.x /= sj; v.y /= sj; } else { v = make_double2(0, 0); }
            if (!f.transposed) f.U[(size_t)dst * m + r] = v;              // U is m x n
            else f.Vh[(size_t)r * n + dst] = make_double2(v.x, -v.y);     // Vh (n x m0=m) row dst = conj(column)
        } else {
            const int rr = r - m;
            if (!pr.V) continue;
            const double2 v = pr.V[(size_t)j * n + rr];
            if (!f.transposed) f.Vh[(size_t)rr * n + dst] = make_double2(v.x, -v.y);  // Vh is n x n
            else f.U[(size_t)dst * n + rr] = v;                                        // U (m0=n x n)
        }
    }
}

// B = A^H (n x m from m x n)
__global__ void conj_transpose_kernel(const double2* __restrict__ A, double2* __restrict__ B, int m, int n) {
    __shared__ double2 t[32][33];
    const int tiles_x = (m + 31) / 32;  // 1-D grid: either dimension may exceed the 65535 limit of grid.y
    const int bx = (int)(blockIdx.x % tiles_x) * 32, by = (int)(blockIdx.x / tiles_x) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = bx + threadIdx.x, c = by + j;
        if (r < m && c < n) t[j][threadIdx.x] = A[(size_t)c * m + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + threadIdx.x, c = bx + j;  // B[r, c] = conj(A[c, r])
        if (r < n && c < m) { const double2 v = t[threadIdx.x][j]; B[(size_t)c * n + r] = make_double2(v.x, -v.y); }
    }
}

// ---------------------------------------------------------------------------------------------
// orthonormal completion of the null vectors of a rank-deficient problem
// ---------------------------------------------------------------------------------------------
// D[i] = pseudo-random in [-1, 1)^2 (splitmix64 of the index: reproducible, no state)
__global__ void fill_random_kernel(double2* __restrict__ D, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)e * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const double a = (double)(z >> 32) * (1.0 / 2147483648.0) - 1.0, b = (double)(z & 0xffffffffull) * (1.0 / 2147483648.0) - 1.0;
        D[e] = make_double2(a, b);
    }
}
__global__ void negate_kernel(double2* __restrict__ P, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = P[e];
        P[e] = make_double2(-v.x, -v.y);
    }
}
// Vh[row0 + j, c] = conj(D[c, j])  (D: len x d column-major; Vh: leading dimension ld)
__global__ void store_conj_rows_kernel(const double2* __restrict__ D, int64_t len, int64_t d, double2* __restrict__ Vh, int64_t ld, int64_t row0) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e % len, j = e / len;
        const double2 v = D[e];
        Vh[c * ld + row0 + j] = make_double2(v.x, -v.y);
    }
}

// The factor built from the normalised Jacobi columns (U, or the rows of Vh when the Jacobi ran on A^H) has `dead`
// zeroed vectors in its last positions.  Replace them: random block, projected twice against the good vectors
// (GEMMs), orthonormalised by CholeskyQR2.  Any orthonormal completion is a valid SVD factor (LAPACK's is as
// arbitrary); the singular values of those positions stay at their (noise-level) computed values.
static int complete_null_vectors(const SvdJob& job, bool transposed, int dead) {
    cudaStream_t st = stream();
    const int64_t r = std::min(job.m0, job.n0), len = transposed ? job.n0 : job.m0, g = r - dead, d = dead;
    PoolBuf D, keep, P, C;
    if (D.alloc((size_t)len * d * 16) || keep.alloc((size_t)len * d * 16) || P.alloc((size_t)std::max<int64_t>(g, 1) * d * 16) ||
        C.alloc((size_t)d * d * 16))
        return fail(QTN_ENOMEM, "svd: null-space completion workspace");
    auto grid = [](int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8)); };
    fill_random_kernel<<<grid(len * d), 256, 0, st>>>((double2*)D.p, len * d);
    count_launch(1);
    int rc;
    for (int pass = 0; pass < 2 && g > 0; ++pass) {
        if (!transposed) {  // good vectors: columns of U (len x g, ld len)
            if ((rc = zgemm_dense('C', 'N', g, d, len, job.U, len, D.p, len, P.p, g, false))) return rc;
            negate_kernel<<<grid(g * d), 256, 0, st>>>((double2*)P.p, g * d);
            if ((rc = zgemm_dense('N', 'N', len, d, g, job.U, len, P.p, g, D.p, len, true))) return rc;
        } else {            // good vectors: conj of the rows of Vh (g x len, ld r)
            if ((rc = zgemm_dense('N', 'N', g, d, len, job.Vh, r, D.p, len, P.p, g, false))) return rc;
            negate_kernel<<<grid(g * d), 256, 0, st>>>((double2*)P.p, g * d);
            if ((rc = zgemm_dense('C', 'N', len, d, g, job.Vh, r, P.p, g, D.p, len, true))) return rc;
        }
        count_launch(1);
    }
    CUDA_TRY(cudaMemcpyAsync(keep.p, D.p, (size_t)len * d * 16, cudaMemcpyDeviceToDevice, st));
    bool ok = false;
    if ((rc = orth_cholqr2(D.p, keep.p, C.p, len, d, &ok))) return rc;
    if (!ok) return fail(QTN_ECUDA, "svd: orthonormal completion of %d null vectors failed", dead);
    if (!transposed) {
        CUDA_TRY(cudaMemcpyAsync((double2*)job.U + (size_t)g * len, D.p, (size_t)len * d * 16, cudaMemcpyDeviceToDevice, st));
    } else {
        store_conj_rows_kernel<<<grid(len * d), 256, 0, st>>>((const double2*)D.p, len, d, (double2*)job.Vh, r, g);
        count_launch(1);
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return QTN_OK;
}

// ---------------------------------------------------------------------------------------------
// host driver (device-resident inputs): batch of problems
// ---------------------------------------------------------------------------------------------
struct SvdWork {
    void* dev = nullptr;
    size_t bytes = 0;
    void* host = nullptr;
    size_t hbytes = 0;
};
static SvdWork g_work;

static int work_reserve(size_t bytes, size_t hbytes) {
    if (bytes > g_work.bytes) {
        cudaStreamSynchronize(stream());
        if (g_work.dev) cudaFree(g_work.dev);
        g_work.bytes = std::max(bytes, (size_t)1 << 22);
        if (cudaMalloc(&g_work.dev, g_work.bytes) != cudaSuccess) { g_work.dev = nullptr; g_work.bytes = 0; return fail(QTN_ENOMEM, "SVD workspace of %zu bytes", bytes); }
    }
    if (hbytes > g_work.hbytes) {
        cudaStreamSynchronize(stream());
        if (g_work.host) cudaFreeHost(g_work.host);
        g_work.hbytes = std::max(hbytes, (size_t)1 << 16);
        if (cudaMallocHost(&g_work.host, g_work.hbytes) != cudaSuccess) { g_work.host = nullptr; g_work.hbytes = 0; return fail(QTN_ENOMEM, "SVD pinned workspace"); }
    }
    return QTN_OK;
}

// Sub-batch streams: independent groups of matrices run their rounds concurrently (created once, never destroyed).
constexpr int kMaxGroups = 4;
static cudaStream_t g_sub[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};    // Gram / update kernels of a sub-batch
static cudaStream_t g_subE[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};   // its eigensolve kernels (high priority)
static cudaEvent_t g_evG[kMaxGroups], g_evE[kMaxGroups];
static cudaEvent_t g_sub_ev = nullptr;

// Every sub-batch has its own stream; its eigensolve runs on a HIGH-priority companion stream, chained by events.
// Measured with the CTA tracer (tools/jacobi_trace.py): a tensor-pipe kernel in flight fills every SM's register file,
// so at equal priority the pending eigensolve CTAs of the other sub-batches wait until it drains, after which all
// eigensolves run together with the tensor pipe idle (26 % of a sweep).  With priority they take the slots the
// running update kernel frees continuously (tensor pipe idle 18 %).  Two alternatives were measured and dropped:
// a per-sweep phase skew between the streams (they fall back into lock-step within a millisecond: convoy effect) and
// an explicit software pipeline with all tensor-pipe kernels on one stream in interleaved sub-batch order (10 %
// slower: every kernel boundary then exposes its partial last wave, which independent streams back-fill).
static int sub_streams_init() {
    if (g_sub_ev) return QTN_OK;
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least, hi = greatest (numerically lower)
    for (int g = 0; g < kMaxGroups; ++g) {
        CUDA_TRY(cudaStreamCreateWithPriority(&g_sub[g], cudaStreamNonBlocking, lo));
        CUDA_TRY(cudaStreamCreateWithPriority(&g_subE[g], cudaStreamNonBlocking, hi));
        CUDA_TRY(cudaEventCreateWithFlags(&g_evG[g], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&g_evE[g], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&g_sub_ev, cudaEventDisableTiming));
    return QTN_OK;
}

// Launch with (pdl) or without the programmatic-stream-serialization attribute.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// One sub-batch: problems [b0, b1) of the (size-sorted) table.
struct SvdGroup {
    int b0 = 0, b1 = 0;
    int max_nb = 2, maxpairs = 1, maxm = 1, maxn = 1;
    bool any_v = false;
    int S = 1, SU = 1;
    int S_tail = 1, SU_tail = 1;   // row splits of the tail sweeps of a dataflow batch (few live pairs: split as for one matrix)
    int tail_sweeps = 0;
    size_t offG = 0, offW = 0, offF = 0;
    bool active = true;
    int sweeps = 0;
    long pairs_per_sweep = 0;   // (matrix, block pair) visits of one sweep
    long prev_active = -1;      // pairs that rotated in the previous sweep (-1: no sweep yet)
    bool cross = false;         // this sweep runs rounds >= 1 with the cross-only Gram kernel
    bool was_cross = false;
    int cross_sweeps = 0;
};

// Runs the batch; k_out / disc_out are host arrays (batch).  Synchronises the library stream.
int svd_batched_device(int batch, const SvdJob* jobs_in, double er, int64_t maxdim, int64_t* k_out, double* disc_out,
                       int* sweeps_out) {
    if (batch <= 0) return QTN_OK;
    cudaStream_t st = stream();
    int rc = sub_streams_init();
    if (rc) return rc;
    // ---- validate; sort by size (columns, then rows, descending) so that sub-batches are homogeneous ----------
    std::vector<int> perm(batch), tr_in(batch);
    for (int b = 0; b < batch; ++b) {
        const int64_t m0 = jobs_in[b].m0, n0 = jobs_in[b].n0;
        if (m0 < 1 || n0 < 1) return fail(QTN_EINVAL, "svd: empty matrix");
        // only the Jacobi column count min(m0, n0) is bounded (block-pair tables are O(n^2 / 256)); the long
        // dimension may be large (MPS(psi): 2 x 2^(M-1); contract_svd_mps: rows grow as 2^j) -- row offsets are 64-bit
        if (std::min(m0, n0) > 32768 || std::max(m0, n0) > ((int64_t)1 << 30) || m0 * n0 > ((int64_t)1 << 33))
            return fail(QTN_EINVAL, "svd: matrix too large (min dimension <= 32768, max dimension <= 2^30, <= 2^33 elements)");
        perm[b] = b;
        tr_in[b] = m0 < n0;
    }
    auto cols = [&](int b) { return std::min(jobs_in[b].m0, jobs_in[b].n0); };
    auto rows = [&](int b) { return std::max(jobs_in[b].m0, jobs_in[b].n0); };
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return cols(x) != cols(y) ? cols(x) > cols(y) : rows(x) > rows(y); });
    std::vector<SvdJob> jobs(batch);
    std::vector<int> tr(batch);
    for (int i = 0; i < batch; ++i) { jobs[i] = jobs_in[perm[i]]; tr[i] = tr_in[perm[i]]; }

    // ---- workspace layout (device): per problem [At (if transposed) m*n] [V n*n] [sig n] [rank n] [stamps] ------
    size_t dev_bytes = 0;
    auto al = [&](size_t bytes) { size_t o = dev_bytes; dev_bytes += (bytes + 255) / 256 * 256; return o; };
    std::vector<size_t> offAt(batch), offV(batch), offSig(batch), offRank(batch), offStamp(batch), offD(batch);
    std::vector<double> work(batch);
    std::vector<size_t> ver_off(batch);
    size_t ver_ints = 0;
    long pairs_per_round = 0;
    int maxn = 0, maxm = 0;
    double total_work = 0;
    for (int b = 0; b < batch; ++b) {
        const int64_t m = rows(perm[b]), n = cols(perm[b]);
        offAt[b] = tr[b] ? al((size_t)m * n * 16) : 0;
        offV[b] = al((size_t)n * n * 16);
        offSig[b] = al((size_t)n * 8);
        offRank[b] = al((size_t)n * 4);
        int nbk = (int)((n + JB - 1) / JB);
        if (nbk & 1) ++nbk;
        offStamp[b] = al((size_t)(nbk + (size_t)nbk * nbk) * 4);
        offD[b] = al((size_t)nbk * JB * JB * 16);
        ver_off[b] = ver_ints;
        ver_ints += (size_t)nbk;
        pairs_per_round += nbk / 2;
        maxn = std::max<int>(maxn, (int)n);
        maxm = std::max<int>(maxm, (int)m);
        work[b] = (double)(m + ((jobs[b].need_v || tr[b]) ? n : 0)) * n * n;  // ~ flops of one sweep
        total_work += work[b];
    }
    // ---- dataflow mode: enough block pairs per round to keep 4 CTAs on every SM busy ----------------------------
    bool flow = pairs_per_round >= 2 * 148 && maxn > 2 * JB;
    if (const char* e = getenv("QTN_JACOBI_FLOW")) { if (atoi(e) == 0) flow = false; else if (atoi(e) == 2) flow = maxn > 2 * JB; }
    const size_t offVer = al((ver_ints + 16) * 4);   // block versions of all problems, then the two task counters
    std::vector<uint2> tasks;
    int ntasks0 = 0;   // tasks of round 0
    if (flow) {
        int max_nb_all = 2;
        for (int b = 0; b < batch; ++b) { int nbk = (int)((cols(perm[b]) + JB - 1) / JB); if (nbk & 1) ++nbk; max_nb_all = std::max(max_nb_all, nbk); }
        for (int round = 0; round < max_nb_all - 1; ++round) {
            for (int b = 0; b < batch; ++b) {
                int nbk = (int)((cols(perm[b]) + JB - 1) / JB);
                if (nbk & 1) ++nbk;
                if (nbk < 2 || round >= nbk - 1) continue;
                for (int pair = 0; pair < nbk / 2; ++pair) tasks.push_back(make_uint2((unsigned)b, ((unsigned)round << 16) | (unsigned)pair));
            }
            if (round == 0) ntasks0 = (int)tasks.size();
        }
    }
    const size_t offTasks = al(tasks.size() * sizeof(uint2) + 16);
    // ---- sub-batches: contiguous in the sorted order, about equal work each -----------------------------------
    int ngroups = flow ? 1 : std::min(batch, kMaxGroups);
    if (const char* e = getenv("QTN_JACOBI_GROUPS")) ngroups = std::max(1, std::min(std::min(batch, kMaxGroups), atoi(e)));
    if (total_work < 4e9 || flow) ngroups = 1;  // small problems: launch-latency-bound, one stream
    std::vector<SvdGroup> groups;
    {
        // cumulative work shares of the sub-batches.  QTN_JACOBI_SKEW=x (default 0 = equal shares) makes them unequal
        // (share of group g proportional to 1 + x g): equal sub-batches have equal round periods and phase-lock.
        double skew = 0.0;
        if (const char* e = getenv("QTN_JACOBI_SKEW")) skew = atof(e);
        std::vector<double> cum(ngroups + 1, 0.0);
        for (int g = 0; g < ngroups; ++g) cum[g + 1] = cum[g] + 1.0 + skew * g;
        for (int g = 0; g <= ngroups; ++g) cum[g] /= cum[ngroups];
        double acc = 0;
        int start = 0;
        for (int b = 0; b < batch; ++b) {
            acc += work[b];
            const int g = (int)groups.size();
            if (b + 1 == batch || (g + 1 < ngroups && acc >= total_work * cum[g + 1] && batch - (b + 1) >= ngroups - (g + 1))) {
                SvdGroup grp;
                grp.b0 = start;
                grp.b1 = b + 1;
                groups.push_back(grp);
                start = b + 1;
            }
        }
    }
    ngroups = (int)groups.size();
    for (auto& grp : groups) {
        long units = 0;  // (matrix, pair) work items of one round
        for (int b = grp.b0; b < grp.b1; ++b) {
            const int m = (int)rows(perm[b]), n = (int)cols(perm[b]);
            int nbk = (n + JB - 1) / JB;
            if (nbk & 1) ++nbk;
            grp.max_nb = std::max(grp.max_nb, nbk);
            grp.maxm = std::max(grp.maxm, m);
            grp.maxn = std::max(grp.maxn, n);
            grp.any_v = grp.any_v || jobs[b].need_v || tr[b];
            units += nbk / 2;
            grp.pairs_per_sweep += (long)(nbk / 2) * (nbk - 1);
        }
        grp.maxpairs = std::max(grp.max_nb / 2, 1);
        // row splits: enough CTAs for ~8 per SM over the whole batch, at least two 16-row (Gram) / 8-row (update)
        // groups per warp
        const long target = std::max<long>(148 * 8 / ngroups, 1);
        const long ggroups = (grp.maxm + 15) / 16, ugroups = (grp.maxm + 7) / 8 + (grp.any_v ? (grp.maxn + 7) / 8 : 0);
        grp.S = (int)std::max<long>(1, std::min<long>((target + units - 1) / units, ggroups / 8));
        grp.SU = (int)std::max<long>(1, std::min<long>((target * (256 / JTHREADS) + units - 1) / units, ugroups / (2 * (JTHREADS / 32))));
        if (const char* e = getenv("QTN_JACOBI_S")) grp.S = std::max(1, atoi(e));
        if (const char* e = getenv("QTN_JACOBI_SU")) grp.SU = std::max(1, atoi(e));
        grp.S_tail = (int)std::max<long>(grp.S, std::min<long>((target + grp.maxpairs - 1) / grp.maxpairs, ggroups / 8));
        grp.SU_tail = (int)std::max<long>(grp.SU, std::min<long>((target * (256 / JTHREADS) + grp.maxpairs - 1) / grp.maxpairs, ugroups / (2 * (JTHREADS / 32))));
        const size_t nb_ = (size_t)(grp.b1 - grp.b0);
        grp.offG = al(nb_ * grp.maxpairs * std::max(grp.S, flow ? grp.S_tail : 1) * GP_ELEMS * 16);   // partial Grams [problem][pair][S][640]
        grp.offW = al(nb_ * grp.maxpairs * JP * JP * 16);            // rotations W   [problem][pair][32*32]
        grp.offF = al(nb_ * grp.maxpairs * 4);                       // apply flags   [problem][pair]
    }
    size_t tab = dev_bytes;
    size_t tab_bytes = (size_t)batch * (sizeof(SvdProblem) + sizeof(FinishArgs) + 2 * sizeof(void*) + 8 + 8 + 8 + 4 + 4 + 4) + 1024;
    dev_bytes += (tab_bytes + 255) / 256 * 256;
    if ((rc = work_reserve(dev_bytes, tab_bytes))) return rc;
    char* base = (char*)g_work.dev;
    // ---- host-side table image ---------------------------------------------------------------------------------
    CUDA_TRY(cudaStreamSynchronize(st));
    char* h = (char*)g_work.host;
    SvdProblem* hp = (SvdProblem*)h;
    FinishArgs* hf = (FinishArgs*)(hp + batch);
    double** hsig = (double**)(hf + batch);
    int** hrank = (int**)(hsig + batch);
    int64_t* hk = (int64_t*)(hrank + batch);
    double* hdisc = (double*)(hk + batch);
    double* hfro = hdisc + batch;
    int* hrot = (int*)(hfro + batch);
    int* hdead = hrot + batch;
    int* hact = hdead + batch;
    char* dtab = base + tab;
    auto dptr = [&](void* hostp) { return dtab + ((char*)hostp - h); };
    for (int b = 0; b < batch; ++b) {
        const int64_t m0 = jobs[b].m0, n0 = jobs[b].n0;
        const int m = (int)(tr[b] ? n0 : m0), n = (int)(tr[b] ? m0 : n0);
        hp[b].A = tr[b] ? (double2*)(base + offAt[b]) : (double2*)jobs[b].A;
        hp[b].V = (jobs[b].need_v || tr[b]) ? (double2*)(base + offV[b]) : nullptr;
        hp[b].m = m;
        hp[b].n = n;
        int nbk = (n + JB - 1) / JB;
        if (nbk & 1) ++nbk;
        hp[b].nblocks = nbk;
        hp[b].last_mod = (int*)(base + offStamp[b]);
        hp[b].last_ok = hp[b].last_mod + nbk;
        hp[b].D = (double2*)(base + offD[b]);
        hp[b].ver = (int*)(base + offVer) + ver_off[b];
        // last_mod = 1, last_ok = 0: every pair starts "modified after its last check"
        CUDA_TRY(cudaMemsetAsync(hp[b].last_ok, 0, (size_t)nbk * nbk * 4, st));
        set_int_kernel<<<(nbk + 255) / 256, 256, 0, st>>>(hp[b].last_mod, nbk, 1);
        hf[b].U = (double2*)jobs[b].U; hf[b].S = jobs[b].S; hf[b].Vh = (double2*)jobs[b].Vh;
        hf[b].k = (int64_t*)dptr(hk + b);
        hf[b].disc = (double*)dptr(hdisc + b);
        hf[b].transposed = tr[b];
        hf[b].m0 = (int)m0; hf[b].n0 = (int)n0;
        hsig[b] = (double*)(base + offSig[b]);
        hrank[b] = (int*)(base + offRank[b]);
        hrot[b] = 0;
        hdead[b] = 0;
        hact[b] = 0;
        hfro[b] = 0.0;
    }
    CUDA_TRY(cudaMemcpyAsync(dtab, h, tab_bytes - 1024, cudaMemcpyHostToDevice, st));
    const SvdProblem* dp = (const SvdProblem*)dtab;
    int* dcounter = (int*)(base + offVer) + ver_ints;
    int* dflow_err = dcounter + 12;
    if (flow) CUDA_TRY(cudaMemsetAsync(dflow_err, 0, 4, st));
    const uint2* dtasks = (const uint2*)(base + offTasks);
    if (flow) {
        CUDA_TRY(cudaMemcpyAsync(base + offTasks, tasks.data(), tasks.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
        static bool carve_done = false;
        if (!carve_done) {  // dynamic shared-memory arenas (45 - 110 KB per CTA)
            CUDA_TRY(cudaFuncSetAttribute(jacobi_flow_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlowCfg<4>::SMEM));
            CUDA_TRY(cudaFuncSetAttribute(jacobi_flow_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlowCfg<3>::SMEM));
            CUDA_TRY(cudaFuncSetAttribute(jacobi_flow_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlowCfg<2>::SMEM));
            CUDA_TRY(cudaFuncSetAttribute(jacobi_flow_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlowCfg<3>::SMEM));
            CUDA_TRY(cudaFuncSetAttribute(jacobi_flow_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlowCfg<2>::SMEM));
            carve_done = true;
        }
    }
    const FinishArgs* df = (const FinishArgs*)dptr(hf);
    double* const* dsig = (double* const*)dptr(hsig);
    int* const* drank = (int* const*)dptr(hrank);
    int* drot = (int*)dptr(hrot);
    int* dact = (int*)dptr(hact);
    double* dfro = (double*)dptr(hfro);
    for (int b = 0; b < batch; ++b)
        if (tr[b]) {
            const unsigned g = (unsigned)(((jobs[b].m0 + 31) / 32) * ((jobs[b].n0 + 31) / 32));
            conj_transpose_kernel<<<g, dim3(32, 8), 0, st>>>((const double2*)jobs[b].A, hp[b].A, (int)jobs[b].m0, (int)jobs[b].n0);
            count_launch(1);
        }
    set_identity_kernel<<<dim3((unsigned)std::min<int64_t>(148 * 4, ((int64_t)maxn * maxn + 255) / 256), batch), 256, 0, st>>>(dp);
    fro_norm_kernel<<<dim3((unsigned)std::min<int64_t>(148, ((int64_t)maxn * maxm + 255) / 256), batch), 256, 0, st>>>(dp, dfro);
    count_launch(2);
    double tol = 1e-15 * std::sqrt((double)std::max(maxm, 1)) * 0.5 + 2.3e-16;
    if (const char* e = getenv("QTN_JACOBI_TOL")) tol *= atof(e);
    int* dstat = nullptr;  // QTN_JACOBI_STATS=1: per-sweep counts of (idle, active) block pairs, printed to stderr
    if (getenv("QTN_JACOBI_STATS")) { cudaMalloc((void**)&dstat, 2 * 64 * sizeof(int)); cudaMemsetAsync(dstat, 0, 2 * 64 * sizeof(int), st); }
    int inner = kInnerSweeps;
    if (const char* e = getenv("QTN_JACOBI_INNER")) inner = std::max(1, atoi(e));
    const bool eig_priority = [] { const char* e = getenv("QTN_JACOBI_EIGPRIO"); return !(e && atoi(e) == 0); }();  // A/B switch
    const bool use_pdl = [] { const char* e = getenv("QTN_JACOBI_PDL"); return !(e && atoi(e) == 0); }();  // A/B switch
    const bool use_cross = [] { const char* e = getenv("QTN_JACOBI_CROSS"); return !(e && atoi(e) == 0); }();  // A/B switch
    // cross-only rotations in the cached-diagonal visits (QTN_JACOBI_XROT=0 / QTN_JACOBI_XROT_FLOW=0 disable them).
    // Three-kernel path: 1536 x 1024 66.3 -> 56.2 ms (15 -> 16 sweeps, 4.42 -> 3.51 ms per sweep), cfg 5 0.383 -> 0.433
    // applies/s; small batches 4 x 512^2 26.1 -> 21.6 ms, 8 x 700 x 520 44.2 -> 38.5 ms, 12 x 256^2 11.3 -> 9.4 ms.  Dataflow kernel: cfg 4 at chi = 512 2.127 -> 2.208 layers/s (22 -> 21 sweeps of the late layers).
    const bool xrot_single = [] { const char* e = getenv("QTN_JACOBI_XROT"); return !(e && atoi(e) == 0); }();
    const bool xrot_flow = [] { const char* e = getenv("QTN_JACOBI_XROT_FLOW"); return !(e && atoi(e) == 0); }();
    // QTN_JACOBI_TAIL=d (A/B switch, default off): sweeps of a dataflow batch that follow one with fewer than 1/d of
    // the pairs rotating run on the three-kernel path with single-matrix row splits.  Measured +1 % on cfg 4 (d = 16:
    // 2.222 -> 2.247 layers/s): the tail sweeps of the task-queue kernel are cheap already.
    const long tail_div = [] { const char* e = getenv("QTN_JACOBI_TAIL"); return e ? atol(e) : 0L; }();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (dstat) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, st); }
    // the sub-batch streams start after the set-up work on the library stream
    CUDA_TRY(cudaEventRecord(g_sub_ev, st));
    for (int g = 0; g < ngroups; ++g) CUDA_TRY(cudaStreamWaitEvent(g_sub[g], g_sub_ev, 0));
    int sweeps = 0, stamp = 1;
    const int kMaxSweeps = 60;
    bool any_active = true;
    const char* trace_file = getenv("QTN_JACOBI_TRACE");
    int trace_sweep = 3;
    if (const char* e = getenv("QTN_JACOBI_TRACE_SWEEP")) trace_sweep = atoi(e);
    unsigned long long* trace_dev = nullptr;
    const unsigned trace_cap = 1u << 20;
    for (; sweeps < kMaxSweeps && any_active; ++sweeps) {
        if (trace_file && sweeps == trace_sweep) {  // record this sweep only
            const unsigned zero = 0;
            CUDA_TRY(cudaMalloc((void**)&trace_dev, (size_t)trace_cap * 24));
            CUDA_TRY(cudaMemcpyToSymbol(d_trace_cap, &trace_cap, sizeof(unsigned)));
            CUDA_TRY(cudaMemcpyToSymbol(d_trace_cnt, &zero, sizeof(unsigned)));
            CUDA_TRY(cudaMemcpyToSymbol(d_trace_buf, &trace_dev, sizeof(void*)));
        }
        // flags: 0 = no rotation yet this sweep, 1 = rotated, -1 = converged (its CTAs exit at once)
        int max_rounds = 0;
        for (int g = 0; g < ngroups; ++g) {
            SvdGroup& grp = groups[g];
            if (!grp.active) continue;
            // Gram mode of this sweep: while most pairs still rotate (the linear-convergence phase) only round 0
            // computes full Grams (which refreshes the per-block cache), the other rounds take the diagonal blocks from
            // the cache.  The polishing sweeps -- and hence every convergence decision that ends the iteration -- use
            // full, freshly computed Grams; pairs accepted during a cached sweep are re-verified (stamps reset).
            grp.cross = use_cross && grp.max_nb > 2 && (grp.prev_active < 0 || grp.prev_active * 4 >= grp.pairs_per_sweep);
            if (grp.was_cross && !grp.cross)
                for (int b = grp.b0; b < grp.b1; ++b)
                    CUDA_TRY(cudaMemsetAsync(hp[b].last_ok, 0, (size_t)hp[b].nblocks * hp[b].nblocks * 4, g_sub[g]));
            grp.was_cross = grp.cross;
            grp.cross_sweeps += grp.cross;
            for (int b = grp.b0; b < grp.b1; ++b) hact[b] = 0;
            CUDA_TRY(cudaMemcpyAsync(drot + grp.b0, hrot + grp.b0, (size_t)(grp.b1 - grp.b0) * 4, cudaMemcpyHostToDevice, g_sub[g]));
            CUDA_TRY(cudaMemcpyAsync(dact + grp.b0, hact + grp.b0, (size_t)(grp.b1 - grp.b0) * 4, cudaMemcpyHostToDevice, g_sub[g]));
            max_rounds = std::max(max_rounds, grp.max_nb - 1);
        }
        int* st_ptr = dstat ? dstat + 2 * std::min(sweeps, 63) : (int*)nullptr;
        // Tail of a dataflow batch: once few pairs still rotate, a sweep of the task-queue kernel is bound by the serial
        // chain of one straggler matrix (63 rounds x one CTA per pair: 7-11 ms whatever the number of live pairs), while
        // three row-split kernels per round walk the same chain in ~4 ms (converged pairs and matrices exit at once).
        const bool tail = flow && tail_div > 0 && groups[0].prev_active >= 0 && groups[0].prev_active * tail_div < groups[0].pairs_per_sweep;
        if (tail) ++groups[0].tail_sweeps;
        if (flow && !tail) {
            // one task-queue kernel per sweep (two when round 0 computes full Grams and the rest take cached diagonals)
            const SvdGroup& grp = groups[0];
            cudaStream_t fs = g_sub[0];
            CUDA_TRY(cudaMemsetAsync(base + offVer, 0, (ver_ints + 8) * 4, fs));   // versions + task counters (not the error flag)
            const int nt = (int)tasks.size();
            static int flow_ctas = -1;   // CTAs per SM of the cross-Gram kernel (QTN_JACOBI_FLOW_CTAS = 2 / 3 / 4)
            if (flow_ctas < 0) { const char* e = getenv("QTN_JACOBI_FLOW_CTAS"); flow_ctas = e ? std::max(2, std::min(4, atoi(e))) : 3; }
            auto launch = [&](bool full, int t0, int t1, int* counter) {
                if (t1 <= t0) return;
                const int per_sm = full ? std::min(flow_ctas, 3) : flow_ctas;
                const int ctas = std::min(t1 - t0, 148 * per_sm);
#define QTN_FLOW_LAUNCH(F, B)                                                                                                   \
    jacobi_flow_kernel<F, B><<<ctas, FLOW_THREADS, FlowCfg<B>::SMEM, fs>>>(dp, dtasks + t0, t1 - t0, counter, tol, drot,          \
                                                                           (const double*)dfro, inner, st_ptr, stamp + 1, dact, dflow_err, (int)xrot_flow)
                if (full) { if (per_sm == 3) QTN_FLOW_LAUNCH(true, 3); else QTN_FLOW_LAUNCH(true, 2); }
                else if (per_sm == 4) QTN_FLOW_LAUNCH(false, 4);
                else if (per_sm == 3) QTN_FLOW_LAUNCH(false, 3);
                else QTN_FLOW_LAUNCH(false, 2);
#undef QTN_FLOW_LAUNCH
                count_launch(1);
            };
            if (grp.cross) { launch(true, 0, ntasks0, dcounter); launch(false, ntasks0, nt, dcounter + 1); }
            else launch(true, 0, nt, dcounter);
            stamp += max_rounds;
            max_rounds = 0;
        }
        for (int round = 0; round < max_rounds; ++round) {
            ++stamp;
            for (int g = 0; g < ngroups; ++g) {  // round-robin over the streams: none is starved by the enqueue order
                const SvdGroup& grp = groups[g];
                if (!grp.active || round >= grp.max_nb - 1) continue;
                const unsigned nb_ = (unsigned)(grp.b1 - grp.b0);
                const int cross = grp.cross && round >= 1;
                // (a companion stream only pays when another sub-batch has tensor-pipe work to overlap with; a single
                // sub-batch chains its three kernels of a round by programmatic dependent launch instead)
                const bool eprio = eig_priority && ngroups > 1;
                const bool pdl = use_pdl && !eprio;
                const int S_cur = tail ? grp.S_tail : grp.S, SU_cur = tail ? grp.SU_tail : grp.SU;
                const dim3 gg((unsigned)(grp.maxpairs * S_cur), nb_);
                if (cross)
                    CUDA_TRY(launch_k(jacobi_gram_cross_kernel, gg, dim3(GRAM_THREADS), g_sub[g], pdl, dp + grp.b0, round, drot + grp.b0, S_cur, grp.maxpairs,
                                      (double2*)(base + grp.offG)));
                else
                    CUDA_TRY(launch_k(jacobi_gram_kernel, gg, dim3(GRAM_THREADS), g_sub[g], pdl, dp + grp.b0, round, drot + grp.b0, S_cur, grp.maxpairs,
                                      (double2*)(base + grp.offG)));
                cudaStream_t se = eprio ? g_subE[g] : g_sub[g];
                if (eprio) { CUDA_TRY(cudaEventRecord(g_evG[g], g_sub[g])); CUDA_TRY(cudaStreamWaitEvent(se, g_evG[g], 0)); }
                // wide eigensolve CTAs for launches that cannot fill the machine anyway (QTN_JACOBI_ET = 128 / 256 forces one)
                {
                    static int et_env = -1;
                    if (et_env < 0) { const char* e = getenv("QTN_JACOBI_ET"); et_env = e ? atoi(e) : 0; }
                    const int et = et_env ? et_env : (((long)grp.maxpairs * nb_ <= 148 || tail) ? 256 : 128);
                    const dim3 eg((unsigned)grp.maxpairs, nb_);
                    if (et == 256)
                        CUDA_TRY(launch_k(jacobi_eig_kernel<256>, eg, dim3(256), se, pdl, dp + grp.b0, round, tol, drot + grp.b0, (const double*)dfro + grp.b0,
                                          inner, st_ptr, stamp, S_cur, grp.maxpairs, (const double2*)(base + grp.offG), (double2*)(base + grp.offW),
                                          (int*)(base + grp.offF), cross, dact + grp.b0, (int)(cross && xrot_single)));
                    else
                        CUDA_TRY(launch_k(jacobi_eig_kernel<128>, eg, dim3(128), se, pdl, dp + grp.b0, round, tol, drot + grp.b0, (const double*)dfro + grp.b0,
                                          inner, st_ptr, stamp, S_cur, grp.maxpairs, (const double2*)(base + grp.offG), (double2*)(base + grp.offW),
                                          (int*)(base + grp.offF), cross, dact + grp.b0, (int)(cross && xrot_single)));
                }
                if (eprio) { CUDA_TRY(cudaEventRecord(g_evE[g], se)); CUDA_TRY(cudaStreamWaitEvent(g_sub[g], g_evE[g], 0)); }
                CUDA_TRY(launch_k(jacobi_update_kernel, dim3((unsigned)(grp.maxpairs * SU_cur), nb_), dim3(JTHREADS), g_sub[g], pdl, dp + grp.b0, round,
                                  drot + grp.b0, SU_cur, grp.maxpairs, (const double2*)(base + grp.offW), (const int*)(base + grp.offF)));
                count_launch(3);
            }
        }
        CUDA_TRY(cudaGetLastError());
        for (int g = 0; g < ngroups; ++g) {
            const SvdGroup& grp = groups[g];
            if (!grp.active) continue;
            CUDA_TRY(cudaMemcpyAsync(hrot + grp.b0, drot + grp.b0, (size_t)(grp.b1 - grp.b0) * 4, cudaMemcpyDeviceToHost, g_sub[g]));
            CUDA_TRY(cudaMemcpyAsync(hact + grp.b0, dact + grp.b0, (size_t)(grp.b1 - grp.b0) * 4, cudaMemcpyDeviceToHost, g_sub[g]));
        }
        for (int g = 0; g < ngroups; ++g)
            if (groups[g].active) CUDA_TRY(cudaStreamSynchronize(g_sub[g]));
        any_active = false;
        if (trace_dev) {  // dump: n records of (kind | smid << 8 | problem << 24, t0, t1) in ns
            unsigned n = 0;
            unsigned long long* nul = nullptr;
            CUDA_TRY(cudaMemcpyFromSymbol(&n, d_trace_cnt, sizeof(unsigned)));
            CUDA_TRY(cudaMemcpyToSymbol(d_trace_buf, &nul, sizeof(void*)));
            n = std::min(n, trace_cap);
            std::vector<unsigned long long> rec((size_t)n * 3);
            CUDA_TRY(cudaMemcpy(rec.data(), trace_dev, (size_t)n * 24, cudaMemcpyDeviceToHost));
            if (FILE* f = fopen(trace_file, "wb")) { fwrite(rec.data(), 24, n, f); fclose(f); }
            cudaFree(trace_dev);
            trace_dev = nullptr;
            trace_file = nullptr;
        }
        for (int g = 0; g < ngroups; ++g) {
            SvdGroup& grp = groups[g];
            if (!grp.active) continue;
            bool any = false;
            grp.prev_active = 0;
            for (int b = grp.b0; b < grp.b1; ++b) {
                grp.prev_active += hact[b];
                // a matrix is only retired by a sweep whose convergence checks all used freshly computed Grams
                if (hrot[b] > 0 || grp.cross) { any = true; hrot[b] = 0; }
                else hrot[b] = -1;
            }
            grp.sweeps = sweeps + 1;
            grp.active = any;
            any_active = any_active || any;
        }
    }
    if (dstat) {
        int hs[128];
        float ms = 0;
        cudaEventRecord(ev1, st);
        cudaEventSynchronize(ev1);
        cudaEventElapsedTime(&ms, ev0, ev1);
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        cudaMemcpy(hs, dstat, sizeof(hs), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[jacobi] batch=%d maxm=%d maxn=%d sweeps=%d %.3f ms (%.3f ms/sweep); groups:", batch, maxm, maxn, sweeps, ms,
                ms / std::max(sweeps, 1));
        for (const auto& grp : groups)
            fprintf(stderr, " [%d..%d) n<=%d m<=%d v=%d S=%d SU=%d sweeps=%d (%d cached-Gram, %d tail);", grp.b0, grp.b1, grp.maxn, grp.maxm, (int)grp.any_v, grp.S,
                    grp.SU, grp.sweeps, grp.cross_sweeps, grp.tail_sweeps);
        fprintf(stderr, " (idle,active) pairs per sweep:");
        for (int i = 0; i < sweeps && i < 64; ++i) fprintf(stderr, " (%d,%d)", hs[2 * i], hs[2 * i + 1]);
        fprintf(stderr, "\n");
        cudaFree(dstat);
    }
    if (flow) {
        int herr = 0;
        CUDA_TRY(cudaMemcpy(&herr, dflow_err, 4, cudaMemcpyDeviceToHost));
        if (herr) return fail(QTN_ECUDA, "Jacobi dataflow kernel: a block-pair task waited for its producers for more than a second (watchdog)");
    }
    // every sub-stream was synchronised by the host above: the library stream continues with the finalisation
    column_norms_kernel<<<dim3(std::min(148 * 2, (maxn + 7) / 8), batch), 256, 0, st>>>(dp, dsig);
    sort_truncate_kernel<<<batch, 256, 0, st>>>(dp, dsig, drank, df, er, (long long)maxdim, (const double*)dfro, (int*)dptr(hdead));
    scatter_factors_kernel<<<dim3(148 * 2, batch), 256, 0, st>>>(dp, dsig, drank, df, (const double*)dfro);
    count_launch(3);
    CUDA_TRY(cudaMemcpyAsync(hk, dptr(hk), (size_t)batch * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(hdead, dptr(hdead), (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    for (int b = 0; b < batch; ++b) {
        if (k_out) k_out[perm[b]] = hk[b];
        if (disc_out) disc_out[perm[b]] = hdisc[b];
    }
    // rank-deficient inputs (rare path): the zeroed null vectors become an orthonormal completion.  The pinned
    // table image is not touched by the GEMMs below, but copy what is needed first: orth_cholqr2 may not reuse it.
    std::vector<int> dead(hdead, hdead + batch);
    for (int b = 0; b < batch; ++b)
        if (dead[b] > 0 && (rc = complete_null_vectors(jobs[b], tr[b] != 0, dead[b]))) return rc;
    if (sweeps_out) *sweeps_out = sweeps;
    if (any_active) return fail(QTN_ECUDA, "Jacobi SVD did not converge in %d sweeps", kMaxSweeps);
    return QTN_OK;
}

}  // namespace qtn

using namespace qtn;

extern "C" {

int qtn_svd_trunc_device(void* dev_a, int64_t m, int64_t n, double er, int64_t maxdim, void* dev_u, double* dev_s,
                         void* dev_vh, int64_t* k_out, int32_t* sweeps_out) {
    QTN_API_GUARD();
    int rc = device_ready();
    if (rc) return rc;
    if (!dev_a || !dev_u || !dev_s || !dev_vh) return fail(QTN_EINVAL, "qtn_svd_trunc_device: null argument");
    SvdJob job{(double2*)dev_a, m, n, (double2*)dev_u, dev_s, (double2*)dev_vh};
    int sw = 0;
    rc = svd_batched_device(1, &job, er, maxdim, k_out, nullptr, &sw);
    if (sweeps_out) *sweeps_out = sw;
    return rc;
}

int qtn_svd_trunc_batched(int32_t batch, const void* const* host_a, const int64_t* m, const int64_t* n, double er,
                          int64_t maxdim, void* const* host_u, double* const* host_s, void* const* host_vh,
                          int64_t* k_out) {
    QTN_API_GUARD();
    int rc = device_ready();
    if (rc) return rc;
    if (batch < 0 || !host_a || !m || !n || !host_u || !host_s || !host_vh) return fail(QTN_EINVAL, "qtn_svd_trunc_batched: null argument");
    if (batch == 0) return QTN_OK;
    cudaStream_t st = stream();
    size_t total = 0;
    std::vector<size_t> oA(batch), oU(batch), oS(batch), oV(batch);
    for (int b = 0; b < batch; ++b) {
        if (m[b] < 1 || n[b] < 1) return fail(QTN_EINVAL, "qtn_svd_trunc: matrix %d is empty", b);
        const int64_t r = std::min(m[b], n[b]);
        auto al = [&](size_t bytes) { size_t o = total; total += (bytes + 255) / 256 * 256; return o; };
        oA[b] = al((size_t)m[b] * n[b] * 16);
        oU[b] = al((size_t)m[b] * r * 16);
        oS[b] = al((size_t)r * 8);
        oV[b] = al((size_t)r * n[b] * 16);
    }
    PoolBuf pool_buf;   // workspace pool: no cudaMalloc / cudaFree (and no implicit device sync) per call
    if (pool_buf.alloc(total)) return QTN_ENOMEM;
    char* buf = (char*)pool_buf.p;
    std::vector<SvdJob> jobs(batch);
    for (int b = 0; b < batch; ++b) {
        cudaMemcpyAsync(buf + oA[b], host_a[b], (size_t)m[b] * n[b] * 16, cudaMemcpyHostToDevice, st);
        jobs[b] = SvdJob{(double2*)(buf + oA[b]), m[b], n[b], (double2*)(buf + oU[b]), (double*)(buf + oS[b]), (double2*)(buf + oV[b])};
    }
    rc = svd_batched_device(batch, jobs.data(), er, maxdim, k_out, nullptr, nullptr);
    if (!rc) {
        for (int b = 0; b < batch; ++b) {
            const int64_t r = std::min(m[b], n[b]);
            cudaMemcpyAsync(host_u[b], buf + oU[b], (size_t)m[b] * r * 16, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(host_s[b], buf + oS[b], (size_t)r * 8, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(host_vh[b], buf + oV[b], (size_t)r * n[b] * 16, cudaMemcpyDeviceToHost, st);
        }
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(QTN_ECUDA, "qtn_svd_trunc: %s", cudaGetErrorString(e));
    }
    return rc;
}

int qtn_svd_trunc(const void* host_a, int64_t m, int64_t n, double er, int64_t maxdim, void* host_u, double* host_s,
                  void* host_vh, int64_t* k_out) {
    QTN_API_GUARD();
    const void* a[1] = {host_a};
    void* u[1] = {host_u};
    double* s[1] = {host_s};
    void* v[1] = {host_vh};
    return qtn_svd_trunc_batched(1, a, &m, &n, er, maxdim, u, s, v, k_out);
}

}  // extern "C"
