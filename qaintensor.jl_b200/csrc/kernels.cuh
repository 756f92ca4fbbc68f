// Device kernels of the contraction path (sm_100a).
//
//  zgemm_gather_kernel   K2: ComplexF64 GEMM on the FP64 tensor pipe
//                        (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no
//                        f64 kind).  Operands are gathered straight from the
//                        tensors' native layouts through additive offset tables
//                        (row offset + k offset), staged to shared memory with
//                        16-byte cp.async in a multi-stage ring, so the TTGT
//                        permutes of the reference (TensorOperations.ncon,
//                        src/contract.jl:257, 263) never touch HBM.
//  zgemm_stream_kernel   K2s: persistent streaming kernel for the tall-skinny steps (M >= 16384, K <= 32, N <= 128):
//                        per-warp cp.async tile rings, A read once, C written once, HBM-bound at 0.9 of the copy peak.
//  zdot_gather_kernel    tiny-M*N / huge-K steps (the closing dot products).
//  permute_gather_kernel / trace_gather_kernel   unary steps.
//  slice_offsets_kernel  per-slice base offsets of the sliced input tensors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qtn {

typedef long long i64;

struct TabArg {
    const i64* lo;
    const i64* hi;
    i64 L;
    int shift;  // log2(L) if L is a power of two, else -1
};

// lo == nullptr: linear mode, offset = i * L (dense matrices need no table).
__device__ __forceinline__ i64 tab_off(const TabArg& t, i64 i) {
    if (t.lo == nullptr) return i * t.L;
    if (t.shift >= 0) return __ldg(t.lo + (i & (t.L - 1))) + __ldg(t.hi + (i >> t.shift));
    i64 q = i / t.L;
    return __ldg(t.lo + (i - q * t.L)) + __ldg(t.hi + q);
}

struct GemmArgs {
    const void* A;  // double2 (ComplexF64) or float2 (ComplexF32 mode) elements
    const void* B;
    void* C;
    const i64* a_soff;  // per-slice base offset (elements) or nullptr
    const i64* b_soff;
    TabArg a_row, a_k, b_k, b_col, c_row, c_col;
    i64 M, N, K;
    i64 k_per_split;
    int tiles_m, tiles_n, group_n;
    int c_dense;
    int mode;  // 0 store, 1 add (single writer), 2 atomic add
    int conj_a, conj_b;  // conjugate the operand on the way into the tensor pipe
    int use_3m;          // contraction plans may use the 3-multiplication complex product (see zgemm_gather_kernel)
    int a_kmajor, b_kmajor;  // operand's contiguous direction is k: consecutive threads of a stage load take consecutive k
    int pdl;                 // launched with programmatic stream serialization (chains of small dense GEMMs: orth.cu, mps.cu)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Element traits of the two precisions (ComplexF64 = double2, ComplexF32 mode = float2).
template <typename T> struct Elem;
template <> struct Elem<double2> { typedef double real; static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); } };
template <> struct Elem<float2> { typedef float real; static __device__ __forceinline__ float2 make(double a, double b) { return make_float2((float)a, (float)b); } };

__device__ __forceinline__ void store_c(float2* p, float re, float im, int mode) {
    if (mode == 0) {
        *p = make_float2(re, im);
    } else if (mode == 1) {
        float2 o = *p;
        *p = make_float2(o.x + re, o.y + im);
    } else {
        atomicAdd(&p->x, re);
        atomicAdd(&p->y, im);
    }
}

__device__ __forceinline__ void store_c(double2* p, double re, double im, int mode) {
    if (mode == 0) {
        *p = make_double2(re, im);
    } else if (mode == 1) {
        double2 o = *p;
        *p = make_double2(o.x + re, o.y + im);
    } else {
        atomicAdd(&p->x, re);
        atomicAdd(&p->y, im);
    }
}

struct UnaryArgs {
    const void* A;
    void* C;
    const i64* a_soff;
    TabArg a_row, a_k, c_row;
    i64 M, K;
    int mode;
};

#ifdef QTN_KERNELS_IMPL
// CTA tile BM x BN complex, warp tile WM x WN, K chunk BK, STAGES-deep cp.async ring.
// K3M = true: complex products by the 3-multiplication (Gauss / "3M") scheme -- P1 = Ar Br,
// P2 = Ai Bi, P3 = (Ar + Ai)(Br + Bi); C = (P1 - P2) + i (P3 - P1 - P2) -- i.e. 3 real DMMAs per
// complex block instead of 4.  Normwise error stays O(eps) (the amplitude bar is 1e-10); the SVD
// path, which needs componentwise accuracy of tiny columns, never uses it.
template <int BM, int BN, int WM, int WN, int BK, int STAGES, int MINB = 1, bool K3M = false>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
zgemm_gather_kernel(const __grid_constant__ GemmArgs g) {
    constexpr int WARPS_M = BM / WM;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int PA = BM + 2, PB = BN + 2;  // smem pitch (complex elements), +32 B kills LDS.128 conflicts
    constexpr int MI = WM / 8, NI = WN / 8;
    constexpr int TPK = NT / BK;  // threads cooperating on one k-row of a stage
    static_assert(NT % BK == 0 && BM % TPK == 0, "tile/thread mismatch");
    constexpr int A_PER = BM / TPK;
    constexpr int B_PER = (BN + TPK - 1) / TPK;

    // programmatic dependent launch: nothing of the predecessor is touched before this (no-op without the launch attribute)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* sA = reinterpret_cast<double2*>(smem_raw);  // [STAGES][BK][PA]
    double2* sB = sA + STAGES * BK * PA;                 // [STAGES][BK][PB]
    i64* sRow = reinterpret_cast<i64*>(sB + STAGES * BK * PB);  // [BM]
    i64* sCol = sRow + BM;                                       // [BN]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    // grouped rasterization: a wave of CTAs covers ~square patches of C, so both operands
    // are re-used out of L2 instead of A being re-streamed once per column tile
    const i64 tile = blockIdx.x;
    const i64 per_group = (i64)g.group_n * g.tiles_m;
    const i64 grp = tile / per_group, in_grp = tile - grp * per_group;
    const i64 gsz = min((i64)g.group_n, (i64)g.tiles_n - grp * g.group_n);
    const i64 m0 = (in_grp / gsz) * BM, n0 = (grp * g.group_n + in_grp % gsz) * BN;
    const i64 kb = (i64)blockIdx.y * g.k_per_split;
    const i64 ke = min(g.K, kb + g.k_per_split);

    const double2* A = static_cast<const double2*>(g.A) + (g.a_soff ? *g.a_soff : 0);
    const double2* B = static_cast<const double2*>(g.B) + (g.b_soff ? *g.b_soff : 0);
    double2* const Cbase = static_cast<double2*>(g.C);

    for (int i = tid; i < BM; i += NT) sRow[i] = (m0 + i < g.M) ? tab_off(g.a_row, m0 + i) : -1;
    for (int i = tid; i < BN; i += NT) sCol[i] = (n0 + i < g.N) ? tab_off(g.b_col, n0 + i) : -1;
    __syncthreads();

    // Stage loads: TPK threads share one k-row.  Consecutive threads walk the rows (columns) of the operand
    // unless its contiguous direction is k (kmajor), where they walk k so that a warp's requests coalesce.
    const int lka = g.a_kmajor ? tid % BK : tid / TPK, lra = g.a_kmajor ? tid / BK : tid % TPK;
    const int lkb = g.b_kmajor ? tid % BK : tid / TPK, lrb = g.b_kmajor ? tid / BK : tid % TPK;
    auto load_stage = [&](int stage, i64 k0) {
        {
            const i64 k = k0 + lka;
            const bool kv = k < ke;
            const i64 ka = kv ? tab_off(g.a_k, k) : 0;
            double2* da = sA + (stage * BK + lka) * PA;
#pragma unroll
            for (int i = 0; i < A_PER; ++i) {
                const int m = lra + i * TPK;
                const i64 ro = sRow[m];
                const bool v = kv && ro >= 0;
                cp_async16(da + m, v ? (A + ro + ka) : A, v);
            }
        }
        {
            const i64 k = k0 + lkb;
            const bool kv = k < ke;
            const i64 kbo = kv ? tab_off(g.b_k, k) : 0;
            double2* db = sB + (stage * BK + lkb) * PB;
#pragma unroll
            for (int i = 0; i < B_PER; ++i) {
                const int n = lrb + i * TPK;
                if (n < BN) {
                    const i64 co = sCol[n];
                    const bool v = kv && co >= 0;
                    cp_async16(db + n, v ? (B + co + kbo) : B, v);
                }
            }
        }
    };

    // 4M: cr = Re, ci = Im.  3M: cr = P1, ci = P2, c3 = P3.
    double cr[MI][NI][2], ci[MI][NI][2], c3[K3M ? MI : 1][K3M ? NI : 1][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
            if (K3M) c3[K3M ? i : 0][K3M ? j : 0][0] = c3[K3M ? i : 0][K3M ? j : 0][1] = 0.0;
        }

    const int nk = (int)((ke - kb + BK - 1) / BK);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, kb + (i64)s * BK);
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<(STAGES >= 2 ? STAGES - 2 : 0)>();
        __syncthreads();
        {
            const int nx = it + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, kb + (i64)nx * BK);
            cp_async_commit();
        }
        const double2* a_s = sA + (it % STAGES) * BK * PA + wm * WM + (lane >> 2);
        const double2* b_s = sB + (it % STAGES) * BK * PB + wn * WN + (lane >> 2);
        // register double-buffered fragments: the LDS.128 of k4+1 fly under the DMMAs of k4
        double2 af[2][MI], bf[2][NI];
        auto load_frags = [&](int buf, int k4) {
            const int kr = k4 * 4 + (lane & 3);
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                af[buf][i] = a_s[kr * PA + i * 8];
                if (g.conj_a) af[buf][i].y = -af[buf][i].y;
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                bf[buf][j] = b_s[kr * PB + j * 8];
                if (g.conj_b) bf[buf][j].y = -bf[buf][j].y;
            }
        };
        load_frags(0, 0);
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            const int cur = k4 & 1;
            if (k4 + 1 < BK / 4) load_frags(cur ^ 1, k4 + 1);
            if constexpr (K3M) {
                double as[MI], bs[NI];
#pragma unroll
                for (int i = 0; i < MI; ++i) as[i] = af[cur][i].x + af[cur][i].y;
#pragma unroll
                for (int j = 0; j < NI; ++j) bs[j] = bf[cur][j].x + bf[cur][j].y;
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], af[cur][i].x, bf[cur][j].x);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], af[cur][i].y, bf[cur][j].y);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(c3[K3M ? i : 0][K3M ? j : 0][0], c3[K3M ? i : 0][K3M ? j : 0][1], as[i], bs[j]);
            } else {
            // four passes of MI*NI independent DMMAs (no back-to-back dependent accumulators)
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], af[cur][i].x, bf[cur][j].x);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], af[cur][i].x, bf[cur][j].y);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], -af[cur][i].y, bf[cur][j].y);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], af[cur][i].y, bf[cur][j].x);
            }
        }
    }
    cp_async_wait<0>();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the dependent grid may fill the SMs this grid's tail frees

    // epilogue: scatter the accumulators through the C offset tables
    i64 crow[MI];
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const i64 r = m0 + wm * WM + i * 8 + (lane >> 2);
        crow[i] = (r < g.M) ? (g.c_dense ? r : tab_off(g.c_row, r)) : -1;
    }
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const i64 c = n0 + wn * WN + j * 8 + 2 * (lane & 3) + q;
            if (c >= g.N) continue;
            const i64 co = g.c_dense ? c * g.M : tab_off(g.c_col, c);
#pragma unroll
            for (int i = 0; i < MI; ++i)
                if (crow[i] >= 0) {
                    double re = cr[i][j][q], im = ci[i][j][q];
                    if (K3M) {
                        const double p1 = re, p2 = im, p3 = c3[K3M ? i : 0][K3M ? j : 0][q];
                        re = p1 - p2;
                        im = p3 - p1 - p2;
                    }
                    store_c(Cbase + crow[i] + co, re, im, g.mode);
                }
        }
}

// K2s: persistent streaming kernel for the tall-skinny steps with a short contraction (N <= 128, K <= 32: "a small
// operator applied to a huge tensor", what a searched / state-vector-like order consists of; HBM-bound for N <= 16).
//   * One CTA per SM, NW independent warps; a warp owns (8 MI)-row tiles  tile = wglobal + i * (gridDim.x * NW)  and
//     its own ring of S stages in shared memory (a stage = one whole A tile, 8 MI rows x KP): loads of the tiles
//     i+1..i+S-1 are in flight (cp.async, 16 B, gathered through the row / k offset tables like the tile kernel) while
//     tile i runs on the FP64 tensor pipe and its C rows are stored straight from the accumulators.  No CTA barrier in
//     the loop (cp.async.wait_group + __syncwarp only): the warps drift apart and load / DMMA / store phases overlap.
//   * B (K x N, at most 64 KB) is staged once per CTA; A is read exactly once, C written exactly once.  N is covered in
//     column blocks of 8 NI (the A fragments are re-read from shared memory per block).
//   * The loop is instruction-lean on purpose (the first version ran 1078 warp instructions per 16 KB tile and was
//     issue-latency-bound at 8 warps per SM): KP is a template parameter so the stage loads unroll with the k offsets
//     pre-scaled to bytes, the dense / full-tile epilogue has no predicates, and small tiles leave room for 16 warps.
//   conj_a is folded into the staging of B and the epilogue: conj(A) B = conj(A conj(B)).
template <int MI, int NI, int KP, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
zgemm_stream_kernel(const __grid_constant__ GemmArgs g, int S) {
    constexpr int TR = MI * 8, PA = TR + 2, CB = NI * 8, NK4 = KP / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NW = blockDim.x >> 5;
    const int K = (int)g.K, N = (int)g.N;
    const int ncb = (N + CB - 1) / CB, PB = ncb * CB + 2;
    double2* sB = reinterpret_cast<double2*>(smem_raw);                 // [KP][PB]
    i64* sKa = reinterpret_cast<i64*>(sB + KP * PB);                     // [KP]  k offsets of A in bytes (-1: padding)
    double2* sA = reinterpret_cast<double2*>(sKa + KP);                  // [NW][S][KP][PA]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, t = lane & 3;
    const char* A = reinterpret_cast<const char*>(static_cast<const double2*>(g.A) + (g.a_soff ? *g.a_soff : 0));
    const double2* B = static_cast<const double2*>(g.B) + (g.b_soff ? *g.b_soff : 0);
    double2* const Cbase = static_cast<double2*>(g.C);
    const bool flip_b = (g.conj_b != 0) != (g.conj_a != 0);
    for (int e = tid; e < KP * ncb * CB; e += blockDim.x) {
        // consecutive threads walk B's contiguous direction
        const int k = g.b_kmajor ? e % KP : e / (ncb * CB), n = g.b_kmajor ? e / KP : e % (ncb * CB);
        double2 v = make_double2(0.0, 0.0);
        if (k < K && n < N) {
            v = __ldg(B + tab_off(g.b_k, k) + tab_off(g.b_col, n));
            if (flip_b) v.y = -v.y;
        }
        sB[k * PB + n] = v;
    }
    for (int k = tid; k < KP; k += blockDim.x) sKa[k] = k < K ? tab_off(g.a_k, k) * 16 : -1;
    __syncthreads();
    const i64 M = g.M;
    const i64 ntiles = (M + TR - 1) / TR;
    const i64 wstride = (i64)gridDim.x * NW;
    const i64 w0 = (i64)blockIdx.x * NW + warp;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(sA + (size_t)warp * S * KP * PA);
    constexpr unsigned STAGE_BYTES = KP * PA * 16;
    // stage loads: lane -> (row, k) with rows fastest, or k fastest when the operand's contiguous direction is k
    constexpr int KPI = 32 / TR;   // k rows per load instruction (rows fastest)
    const int lrow = lane % TR, lksub = lane / TR;
    const unsigned ldst = (unsigned)((lksub * PA + lrow) * 16);
    constexpr int KSH = KP >= 32 ? 5 : (KP >= 16 ? 4 : (KP >= 8 ? 3 : 2));
    const int kk = lane & (KP - 1), rsub = lane >> KSH;   // (k fastest; KP = 32 -> one row per instruction)
    constexpr int RPER = 32 >> KSH;
    auto issue = [&](i64 tile, int stage) {
        const unsigned dst = ring_s + (unsigned)stage * STAGE_BYTES;
        const i64 r = tile * TR + lrow;
        const i64 ro = (tile < ntiles && r < M) ? tab_off(g.a_row, r) * 16 : -1;   // lanes >= TR duplicate (TR = 16)
        if (!g.a_kmajor) {
            const char* src = A + ro;
#pragma unroll
            for (int k = 0; k < KP; k += KPI) {
                const i64 ka = sKa[k + lksub];
                const int sz = (ro | ka) >= 0 ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + ldst + (unsigned)(k * PA * 16)), "l"(sz ? src + ka : A), "r"(sz));
            }
        } else {
            const i64 ka = sKa[kk];
#pragma unroll
            for (int r0 = 0; r0 < TR; r0 += RPER) {
                const i64 rr = __shfl_sync(0xffffffffu, ro, r0 + rsub);
                const int sz = (rr | ka) >= 0 ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + (unsigned)((kk * PA + r0 + rsub) * 16)), "l"(sz ? A + rr + ka : A), "r"(sz));
            }
        }
    };
    for (int s = 0; s < S - 1; ++s) {
        issue(w0 + (i64)s * wstride, s);
        cp_async_commit();
    }
    const bool fast_cols = g.c_dense && g.mode == 0 && N == ncb * CB;
    // N = 4 in a single 8-column block (a two-qubit gate: most steps of a default-order walk): lanes t < 2 hold the four
    // real columns and store them through the same predicate-free path
    const bool half_cols = NI == 1 && g.c_dense && g.mode == 0 && N == 4;
    int stage = 0;
    for (i64 tile = w0; tile < ntiles; tile += wstride) {
        // S - 2 groups may stay pending: the oldest (this tile's) has landed for this thread; __syncwarp makes the
        // other lanes' copies visible and guarantees every lane is done reading the stage overwritten next
        if (S >= 4) cp_async_wait<2>(); else if (S == 3) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        {
            int ns = stage + S - 1;
            if (ns >= S) ns -= S;
            issue(tile + (i64)(S - 1) * wstride, ns);
            cp_async_commit();
        }
        const double2* a_s = sA + ((size_t)warp * S + stage) * (KP * PA) + gq;
        const i64 rbase = tile * TR + gq;
        const bool fast = (fast_cols || half_cols) && (tile + 1) * TR <= M;
        for (int cb = 0; cb < ncb; ++cb) {
            const double2* b_s = sB + cb * CB + gq;
            double cr[MI][NI][2], ci[MI][NI][2];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
            double2 af0[MI], bf0[NI], af1[MI], bf1[NI];   // two statically named fragment buffers (no local-memory indexing)
            auto load_frags = [&](double2 (&af)[MI], double2 (&bf)[NI], int k4) {
                const int kr = k4 * 4 + t;
#pragma unroll
                for (int i = 0; i < MI; ++i) af[i] = a_s[kr * PA + i * 8];
#pragma unroll
                for (int j = 0; j < NI; ++j) bf[j] = b_s[kr * PB + j * 8];
            };
            auto mma = [&](const double2 (&af)[MI], const double2 (&bf)[NI]) {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], af[i].x, bf[j].x);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], af[i].x, bf[j].y);
                double nby[NI];
#pragma unroll
                for (int j = 0; j < NI; ++j) nby[j] = -bf[j].y;
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(cr[i][j][0], cr[i][j][1], af[i].y, nby[j]);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(ci[i][j][0], ci[i][j][1], af[i].y, bf[j].x);
            };
            load_frags(af0, bf0, 0);
            if constexpr (NK4 == 1) {   // K <= 4 (a two-qubit gate on a big tensor): one k4 step, nothing to prefetch
                mma(af0, bf0);
            } else {
#pragma unroll
                for (int k4 = 0; k4 < NK4; k4 += 2) {   // the LDS.128 of the next k4 fly under the DMMAs of the current one
                    load_frags(af1, bf1, k4 + 1);
                    mma(af0, bf0);
                    if (k4 + 2 < NK4) load_frags(af0, bf0, k4 + 2);
                    mma(af1, bf1);
                }
            }
            if (g.conj_a) {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) { ci[i][j][0] = -ci[i][j][0]; ci[i][j][1] = -ci[i][j][1]; }
            }
            // epilogue: straight from the accumulators (8 consecutive rows per column = one full 128-byte line per store)
            if (fast) {
                if (half_cols && t >= 2) continue;   // ncb == 1: nothing else to do for this tile
                double2* cp = Cbase + rbase + (i64)(cb * CB + 2 * t) * M;
#pragma unroll
                for (int j = 0; j < NI; ++j) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
#pragma unroll
                        for (int i = 0; i < MI; ++i) cp[i * 8] = make_double2(cr[i][j][q], ci[i][j][q]);
                        cp += M;
                    }
                    cp += 6 * M;
                }
            } else {
                i64 crow[MI];
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const i64 r = rbase + i * 8;
                    crow[i] = (r < M) ? (g.c_dense ? r : tab_off(g.c_row, r)) : -1;
                }
#pragma unroll
                for (int j = 0; j < NI; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int c = cb * CB + j * 8 + 2 * t + q;
                        if (c >= N) continue;
                        const i64 co = g.c_dense ? (i64)c * M : tab_off(g.c_col, c);
#pragma unroll
                        for (int i = 0; i < MI; ++i)
                            if (crow[i] >= 0) store_c(Cbase + crow[i] + co, cr[i][j][q], ci[i][j][q], g.mode);
                    }
            }
        }
        if (++stage == S) stage = 0;
    }
    cp_async_wait<0>();
}

// K3: ComplexF32 mode on the FP32 pipes (optional precision; tensor cores would need TF32, whose
// 10-bit mantissa misses the 1e-4 amplitude bar).  Same gather/scatter addressing as the FP64
// kernel; 64x64 CTA tile, 256 threads, 4x4 complex outputs per thread, BK = 16, 2-stage cp.async (8 B).
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}

__global__ void __launch_bounds__(256, 2) cgemm_gather_kernel(const __grid_constant__ GemmArgs g) {
    // CTA tile 128 x 64, thread tile 8 x 4 complex (rows tx + 16 i, columns ty * 4 + j), BK = 8, 3 stages, 2 CTAs/SM
    // (measured on the cfg-3 dominant step: 38.2 TFLOP/s; 1 CTA/SM 32.4; 128 threads with 8 x 8 thread tiles 34.3)
    constexpr int BM = 128, BN = 64, BK = 8, NT = 256, PA = BM + 2, PB = BN + 2, STAGES = 3, TM = 8, TN = 4;
    __shared__ __align__(16) float2 sA[STAGES][BK][PA];
    __shared__ __align__(16) float2 sB[STAGES][BK][PB];
    __shared__ i64 sRow[BM], sCol[BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const i64 tile = blockIdx.x;
    const i64 per_group = (i64)g.group_n * g.tiles_m;
    const i64 grp = tile / per_group, in_grp = tile - grp * per_group;
    const i64 gsz = min((i64)g.group_n, (i64)g.tiles_n - grp * g.group_n);
    const i64 m0 = (in_grp / gsz) * BM, n0 = (grp * g.group_n + in_grp % gsz) * BN;
    const i64 kb = (i64)blockIdx.y * g.k_per_split;
    const i64 ke = min(g.K, kb + g.k_per_split);
    const float2* A = static_cast<const float2*>(g.A) + (g.a_soff ? *g.a_soff : 0);
    const float2* B = static_cast<const float2*>(g.B) + (g.b_soff ? *g.b_soff : 0);
    float2* const Cbase = static_cast<float2*>(g.C);
    for (int i = tid; i < BM; i += NT) sRow[i] = (m0 + i < g.M) ? tab_off(g.a_row, m0 + i) : -1;
    for (int i = tid; i < BN; i += NT) sCol[i] = (n0 + i < g.N) ? tab_off(g.b_col, n0 + i) : -1;
    __syncthreads();
    const int lk = tid >> 5, lr = tid & 31;  // 32 threads per k-row of a stage
    auto load_stage = [&](int stage, i64 k0) {
        const i64 k = k0 + lk;
        const bool kv = k < ke;
        const i64 ka = kv ? tab_off(g.a_k, k) : 0, kbo = kv ? tab_off(g.b_k, k) : 0;
#pragma unroll
        for (int i = 0; i < BM / 32; ++i) {
            const int m = lr + 32 * i;
            const i64 ro = sRow[m];
            const bool v = kv && ro >= 0;
            cp_async8(&sA[stage][lk][m], v ? (A + ro + ka) : A, v);
        }
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
            const int n = lr + 32 * i;
            const i64 co = sCol[n];
            const bool v = kv && co >= 0;
            cp_async8(&sB[stage][lk][n], v ? (B + co + kbo) : B, v);
        }
    };
    float cr[TM][TN], ci[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) cr[i][j] = ci[i][j] = 0.f;
    const float sa = g.conj_a ? -1.f : 1.f, sb = g.conj_b ? -1.f : 1.f;
    const bool conj_any = g.conj_a || g.conj_b;
    const int nk = (int)((ke - kb + BK - 1) / BK);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, kb + (i64)s * BK);
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = it + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, kb + (i64)nx * BK);
            cp_async_commit();
        }
        const int st = it % STAGES;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float2 a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = sA[st][k][tx + 16 * i];
            const float4 b01 = *reinterpret_cast<const float4*>(&sB[st][k][ty * 4]);
            const float4 b23 = *reinterpret_cast<const float4*>(&sB[st][k][ty * 4 + 2]);
            b[0] = make_float2(b01.x, b01.y); b[1] = make_float2(b01.z, b01.w);
            b[2] = make_float2(b23.x, b23.y); b[3] = make_float2(b23.z, b23.w);
            if (conj_any) {  // uniform, rare (dense 'C' operands only)
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i].y *= sa;
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j].y *= sb;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    cr[i][j] = fmaf(a[i].x, b[j].x, cr[i][j]);
                    cr[i][j] = fmaf(-a[i].y, b[j].y, cr[i][j]);
                    ci[i][j] = fmaf(a[i].x, b[j].y, ci[i][j]);
                    ci[i][j] = fmaf(a[i].y, b[j].x, ci[i][j]);
                }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const i64 c = n0 + ty * 4 + j;
        if (c >= g.N) continue;
        const i64 co = g.c_dense ? c * g.M : tab_off(g.c_col, c);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const i64 r = m0 + tx + 16 * i;
            if (r >= g.M) continue;
            const i64 ro = g.c_dense ? r : tab_off(g.c_row, r);
            store_c(Cbase + ro + co, cr[i][j], ci[i][j], g.mode);
        }
    }
}

// C[m, n] += sum_k A[m, k] B[k, n] for M*N <= 16 and huge K (the closing dot products):
// every thread strides over k, block-reduces, one atomic per (m, n) per CTA.  C must
// be pre-zeroed (or hold the running sum when accumulating).
template <typename T>
__global__ void __launch_bounds__(256) zdot_gather_kernel(const __grid_constant__ GemmArgs g) {
    typedef typename Elem<T>::real real;
    const T* A = static_cast<const T*>(g.A) + (g.a_soff ? *g.a_soff : 0);
    const T* B = static_cast<const T*>(g.B) + (g.b_soff ? *g.b_soff : 0);
    const int M = (int)g.M, N = (int)g.N;
    __shared__ double red[8][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = 0; e < M * N; ++e) {
        const int m = e % M, n = e / M;
        const i64 ro = tab_off(g.a_row, m), co = tab_off(g.b_col, n);
        double r0 = 0, i0 = 0, r1 = 0, i1 = 0;
        const i64 stride = (i64)gridDim.x * blockDim.x;
        i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
        for (; k + stride < g.K; k += 2 * stride) {
            const T a0 = __ldg(A + ro + tab_off(g.a_k, k)), b0 = __ldg(B + co + tab_off(g.b_k, k));
            const T a1 = __ldg(A + ro + tab_off(g.a_k, k + stride)), b1 = __ldg(B + co + tab_off(g.b_k, k + stride));
            r0 += (double)a0.x * b0.x - (double)a0.y * b0.y;   // always accumulated in double
            i0 += (double)a0.x * b0.y + (double)a0.y * b0.x;
            r1 += (double)a1.x * b1.x - (double)a1.y * b1.y;
            i1 += (double)a1.x * b1.y + (double)a1.y * b1.x;
        }
        if (k < g.K) {
            const T a0 = __ldg(A + ro + tab_off(g.a_k, k)), b0 = __ldg(B + co + tab_off(g.b_k, k));
            r0 += (double)a0.x * b0.x - (double)a0.y * b0.y;
            i0 += (double)a0.x * b0.y + (double)a0.y * b0.x;
        }
        double r = r0 + r1, i = i0 + i1;
        for (int o = 16; o > 0; o >>= 1) {
            r += __shfl_xor_sync(0xffffffffu, r, o);
            i += __shfl_xor_sync(0xffffffffu, i, o);
        }
        __syncthreads();
        if (lane == 0) { red[warp][0] = r; red[warp][1] = i; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double s = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
            const i64 off = g.c_dense ? (m + (i64)n * g.M) : (tab_off(g.c_row, m) + tab_off(g.c_col, n));
            atomicAdd(reinterpret_cast<real*>(static_cast<T*>(g.C) + off) + threadIdx.x, (real)s);
        }
    }
}


// out[c_row(i)] (+)= in[a_row(i)]; i enumerates the OUTPUT layout (coalesced writes).
template <typename T>
__global__ void __launch_bounds__(256) permute_gather_kernel(const __grid_constant__ UnaryArgs g) {
    const T* A = static_cast<const T*>(g.A) + (g.a_soff ? *g.a_soff : 0);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < g.M; i += (i64)gridDim.x * blockDim.x) {
        const T v = __ldg(A + tab_off(g.a_row, i));
        store_c(static_cast<T*>(g.C) + tab_off(g.c_row, i), v.x, v.y, g.mode);
    }
}

// out[i] = sum_t in[a_row(i) + a_k(t)]  (partial trace, src/network2graph.jl:436-445 networks)
template <typename T>
__global__ void __launch_bounds__(256) trace_gather_kernel(const __grid_constant__ UnaryArgs g) {
    const T* A = static_cast<const T*>(g.A) + (g.a_soff ? *g.a_soff : 0);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < g.M; i += (i64)gridDim.x * blockDim.x) {
        const i64 ro = tab_off(g.a_row, i);
        double re = 0, im = 0;
        for (i64 t = 0; t < g.K; ++t) {
            const T v = __ldg(A + ro + tab_off(g.a_k, t));
            re += v.x;
            im += v.y;
        }
        static_cast<T*>(g.C)[i] = Elem<T>::make(re, im);
    }
}

// soff[t] = sum over the sliced labels on input t of digit(label) * stride; then sid += 1.
__global__ void slice_offsets_kernel(i64* sid_ptr, int nlab, const i64* slice_dims, int nt,
                                     const int* first, const int* pos, const i64* stride, i64* soff) {
    const i64 sid = *sid_ptr;
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        i64 off = 0;
        for (int e = first[t]; e < first[t + 1]; ++e) {
            i64 r = sid;
            for (int q = 0; q < pos[e]; ++q) r /= slice_dims[q];
            off += (r % slice_dims[pos[e]]) * stride[e];
        }
        soff[t] = off;
    }
    __syncthreads();
    if (threadIdx.x == 0) *sid_ptr = sid + 1;
}

#endif  // QTN_KERNELS_IMPL

}  // namespace qtn
