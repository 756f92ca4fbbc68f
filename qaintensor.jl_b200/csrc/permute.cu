// K1: tensor permute (Julia `permutedims`, src/contract.jl:244, src/svd.jl:20-21,
// src/switch.jl:29-35) as a shared-memory-staged, 128-bit-per-element transpose.
//
// A ComplexF64 element is exactly one 16-byte vector.  After dropping extent-1 modes
// and fusing modes that stay adjacent, a tile is the set T of modes made of the
// input-fastest modes (contiguous reads) plus the output-fastest modes (contiguous
// writes).  A CTA reads its tile in input order, parks it in (padded) shared memory
// and writes it back in output order; the remaining modes are enumerated by the grid
// through two-level additive offset tables (same scheme as the GEMM operands).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "qtn_internal.h"

namespace qtn {
cudaStream_t stream();
void count_launch(int64_t n);

struct PermArgs {
    const double2* in;
    double2* out;
    const i64* tile_in;     // [TT] input offset of tile element e (input-order enumeration)
    const i64* tile_out;    // [TT] output offset of tile element e' (output-order enumeration)
    const int* tile_slot;   // [TT] padded smem slot of element e' (its input-order position)
    TabArg rest_in, rest_out;
    i64 nrest;
    int TT;
};

__device__ __forceinline__ int pad_slot(int s) { return s + (s >> 5); }

__global__ void __launch_bounds__(256) permute_tiled_kernel(const __grid_constant__ PermArgs g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    for (i64 r = blockIdx.x; r < g.nrest; r += gridDim.x) {
        const double2* src = g.in + tab_off(g.rest_in, r);
        double2* dst = g.out + tab_off(g.rest_out, r);
        for (int e = threadIdx.x; e < g.TT; e += blockDim.x) {
            double2 v;
            const double2* p = src + __ldg(g.tile_in + e);
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
            tile[pad_slot(e)] = v;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < g.TT; e += blockDim.x) {
            const double2 v = tile[__ldg(g.tile_slot + e)];
            double2* p = dst + __ldg(g.tile_out + e);
            asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = in[i];
}

// device scratch for the per-call tables (stream-ordered reuse)
static void* g_scratch = nullptr;
static size_t g_scratch_bytes = 0;
static void* g_hscratch = nullptr;

static int scratch_reserve(size_t bytes) {
    if (bytes <= g_scratch_bytes) return QTN_OK;
    cudaStreamSynchronize(stream());
    if (g_scratch) cudaFree(g_scratch);
    if (g_hscratch) cudaFreeHost(g_hscratch);
    size_t nb = std::max<size_t>(bytes, 1 << 20);
    if (cudaMalloc(&g_scratch, nb) != cudaSuccess || cudaMallocHost(&g_hscratch, nb) != cudaSuccess) {
        g_scratch = nullptr; g_hscratch = nullptr; g_scratch_bytes = 0;
        return fail(QTN_ENOMEM, "permute scratch allocation of %zu bytes failed", nb);
    }
    g_scratch_bytes = nb;
    return QTN_OK;
}

struct Mode {
    int64_t ext, sin, sout;
};

int permutedims_device(const void* in, int rank, const int64_t* dims, const int32_t* perm, void* out) {
    if (rank < 0 || rank > 64) return fail(QTN_EINVAL, "permutedims: unsupported rank %d", rank);
    std::vector<int> seen(rank, 0);
    for (int i = 0; i < rank; ++i) {
        if (perm[i] < 1 || perm[i] > rank || seen[perm[i] - 1]) return fail(QTN_EINVAL, "permutedims: perm is not a permutation of 1..%d", rank);
        seen[perm[i] - 1] = 1;
    }
    int64_t total = 1;
    for (int i = 0; i < rank; ++i) { if (dims[i] < 0) return fail(QTN_EINVAL, "permutedims: negative extent"); total *= dims[i]; }
    if (total == 0) return QTN_OK;
    cudaStream_t st = stream();
    // modes in OUTPUT order with their input strides
    std::vector<int64_t> sin_all(rank);
    { int64_t s = 1; for (int i = 0; i < rank; ++i) { sin_all[i] = s; s *= dims[i]; } }
    std::vector<Mode> om;  // output order
    for (int i = 0; i < rank; ++i) {
        int a = perm[i] - 1;
        if (dims[a] == 1) continue;
        if (!om.empty() && om.back().sin * om.back().ext == sin_all[a]) om.back().ext *= dims[a];  // fuse
        else om.push_back({dims[a], sin_all[a], 0});
    }
    { int64_t s = 1; for (auto& m : om) { m.sout = s; s *= m.ext; } }
    if (om.size() <= 1) {
        int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
        copy_kernel<<<blocks, 256, 0, st>>>((const double2*)in, (double2*)out, total);
        count_launch(1);
        return cudaGetLastError() == cudaSuccess ? QTN_OK : fail(QTN_ECUDA, "copy kernel launch failed");
    }
    // split big modes into factors <= 32 so tiles can take chunks of them (4096 -> 32 x 32 x 4)
    std::vector<Mode> modes;
    for (auto& m : om) {
        int64_t ext = m.ext, sin = m.sin, sout = m.sout;
        while (ext > 32) {
            int64_t c = 1;
            for (int64_t d = 32; d >= 2; --d) if (ext % d == 0) { c = d; break; }
            if (c == 1) break;  // large prime factor: keep the remainder whole
            modes.push_back({c, sin, sout});
            ext /= c; sin *= c; sout *= c;
        }
        modes.push_back({ext, sin, sout});
    }
    int nm = (int)modes.size();
    std::vector<int> by_in(nm), by_out(nm);
    for (int i = 0; i < nm; ++i) by_in[i] = by_out[i] = i;
    std::sort(by_in.begin(), by_in.end(), [&](int a, int b) { return modes[a].sin < modes[b].sin; });
    std::sort(by_out.begin(), by_out.end(), [&](int a, int b) { return modes[a].sout < modes[b].sout; });
    std::vector<char> inT(nm, 0);
    int64_t TT = 1;
    const int64_t kMaxTile = 4096;
    {
        int64_t run = 1;
        for (int i : by_in) { if (run >= 32 || TT * modes[i].ext > kMaxTile) break; inT[i] = 1; TT *= modes[i].ext; run *= modes[i].ext; }
        run = 1;
        for (int i : by_out) {
            if (run >= 32) break;
            if (!inT[i]) { if (TT * modes[i].ext > kMaxTile) break; inT[i] = 1; TT *= modes[i].ext; }
            run *= modes[i].ext;
        }
        for (int i : by_in) { if (TT >= 1024) break; if (!inT[i] && TT * modes[i].ext <= 2048) { inT[i] = 1; TT *= modes[i].ext; } }
    }
    std::vector<int> t_in, t_out;  // tile modes in input / output order
    for (int i : by_in) if (inT[i]) t_in.push_back(i);
    for (int i : by_out) if (inT[i]) t_out.push_back(i);
    std::vector<int64_t> re, rsi, rso;
    for (int i : by_out) if (!inT[i]) { re.push_back(modes[i].ext); rsi.push_back(modes[i].sin); rso.push_back(modes[i].sout); }
    int64_t nrest = 1;
    for (auto e : re) nrest *= e;

    std::vector<int64_t> tabs;
    tabs.push_back(0);
    OffTable tr_in = make_table(tabs, re, rsi, 4096), tr_out = make_table(tabs, re, rso, 4096);
    int64_t pos_tile_in = (int64_t)tabs.size();
    tabs.resize(tabs.size() + 2 * TT);
    std::vector<int> slots(TT);
    // input-order enumeration: position weight of each tile mode
    std::vector<int64_t> w_in(nm, 0);
    { int64_t w = 1; for (int i : t_in) { w_in[i] = w; w *= modes[i].ext; } }
    for (int64_t e = 0; e < TT; ++e) {
        int64_t r = e, off = 0;
        for (int i : t_in) { off += (r % modes[i].ext) * modes[i].sin; r /= modes[i].ext; }
        tabs[pos_tile_in + e] = off;
    }
    for (int64_t e = 0; e < TT; ++e) {
        int64_t r = e, off = 0, slot = 0;
        for (int i : t_out) { int64_t d = r % modes[i].ext; off += d * modes[i].sout; slot += d * w_in[i]; r /= modes[i].ext; }
        tabs[pos_tile_in + TT + e] = off;
        slots[e] = (int)(slot + (slot >> 5));
    }
    size_t bytes = tabs.size() * 8 + (size_t)TT * 4;
    int rc = scratch_reserve(bytes);
    if (rc) return rc;
    cudaStreamSynchronize(st);  // host staging buffer reuse
    memcpy(g_hscratch, tabs.data(), tabs.size() * 8);
    memcpy((char*)g_hscratch + tabs.size() * 8, slots.data(), (size_t)TT * 4);
    if (cudaMemcpyAsync(g_scratch, g_hscratch, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail(QTN_ECUDA, "permutedims: table upload failed");
    const i64* dt = (const i64*)g_scratch;
    auto targ = [&](const OffTable& t) {
        TabArg a; a.lo = dt + t.lo; a.hi = dt + t.hi; a.L = t.L; a.shift = -1;
        if ((t.L & (t.L - 1)) == 0) { int s = 0; while (((i64)1 << s) < t.L) ++s; a.shift = s; }
        return a;
    };
    PermArgs g;
    g.in = (const double2*)in;
    g.out = (double2*)out;
    g.tile_in = dt + pos_tile_in;
    g.tile_out = dt + pos_tile_in + TT;
    g.tile_slot = (const int*)((const char*)g_scratch + tabs.size() * 8);
    g.rest_in = targ(tr_in);
    g.rest_out = targ(tr_out);
    g.nrest = nrest;
    g.TT = (int)TT;
    size_t smem = (size_t)(TT + (TT >> 5) + 1) * 16;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(permute_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024); attr = true; }
    int blocks = (int)std::min<int64_t>(nrest, 148 * 8);
    permute_tiled_kernel<<<blocks, 256, smem, st>>>(g);
    if (cudaGetLastError() != cudaSuccess) return fail(QTN_ECUDA, "permute kernel launch failed");
    count_launch(1);
    return QTN_OK;
}

}  // namespace qtn
