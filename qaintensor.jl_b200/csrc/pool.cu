// Grow-only device workspace pool: the transient buffers of the SVD / MPS / MPO entry points
// (qtn_svd_trunc[_batched], the null-vector completion, CholeskyQR2, qtn_contract_svd, MPO apply / expectation,
// permutedims) come from here instead of cudaMalloc / cudaFree per call, so a steady-state caller -- decompose!,
// MPO(m), switch!, the MPS sweeps: one SVD after another -- pays no allocation and no implicit device
// synchronisation (cudaFree synchronises).  Blocks are cached by size and re-used best-fit; nothing is returned
// to the driver before qtn_shutdown (or an allocation failure, which first drops the cache and retries).
//
// The device state of contraction plans (inputs, arena, offset tables; exec.cu) comes from here as well when a buffer is
// at most kPoolMaxBlock: a one-shot `contract(net)` -- plan, upload, execute, destroy -- then costs no cudaMalloc /
// cudaFree / cudaMallocHost either (they were 60 % of the call on the reference's own QFT-20 benchmark network).
// Larger buffers (the 100 GB arenas of sliced contractions) are allocated and freed directly, and the cache of idle
// blocks is capped (kPoolMaxIdle): beyond it a freed block goes back to the driver.  pinned_alloc / pinned_free are the
// same cache for page-locked host staging buffers.
//
// Safety of re-use: every kernel of the library that touches a pool block is ordered on the library stream (the SVD's
// sub-batch streams are joined by the host before svd_batched_device returns), so a block handed out again is
// only written after the previous user's work in stream order.  Not re-entrant, like the rest of the library.
#include <cuda_runtime.h>

#include <map>
#include <unordered_map>

#include "qtn_internal.h"

namespace qtn {
namespace {
std::multimap<size_t, void*> g_free;           // cached blocks by capacity
std::unordered_map<void*, size_t> g_live;      // capacity of every block handed out
int64_t g_mallocs = 0, g_hits = 0;
size_t g_bytes = 0;                            // total capacity owned (live + cached)
size_t g_idle = 0;                             // capacity of the cached (idle) blocks
const size_t kPoolMaxIdle = (size_t)8 << 30;   // idle blocks beyond this go back to the driver

std::multimap<size_t, void*> g_pfree;          // page-locked host blocks, same scheme
std::unordered_map<void*, size_t> g_plive;
size_t g_pidle = 0;
const size_t kPinnedMaxIdle = (size_t)256 << 20;

size_t round_up(size_t bytes) {
    if (bytes < 256) return 256;
    if (bytes <= ((size_t)1 << 20)) return (bytes + 255) / 256 * 256;
    return (bytes + (((size_t)1 << 20) - 1)) >> 20 << 20;   // 1 MiB granules: similar requests share blocks
}
}  // namespace

void* pool_alloc(size_t bytes) {
    const size_t need = round_up(bytes);
    auto it = g_free.lower_bound(need);
    // a cached block serves the request unless it is grossly larger (keep the big blocks for big requests)
    if (it != g_free.end() && (it->first <= 2 * need || it->first <= ((size_t)4 << 20))) {
        void* p = it->second;
        g_live[p] = it->first;
        g_idle -= it->first;
        g_free.erase(it);
        ++g_hits;
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) {   // drop the cache and try once more
        cudaGetLastError();
        pool_trim();
        e = cudaMalloc(&p, need);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(QTN_ENOMEM, "device workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    ++g_mallocs;
    g_bytes += need;
    g_live[p] = need;
    return p;
}

void pool_free(void* p) {
    if (!p) return;
    auto it = g_live.find(p);
    if (it == g_live.end()) return;   // not ours
    if (g_idle + it->second > kPoolMaxIdle) {   // cache full: back to the driver (cudaFree synchronises the device)
        cudaFree(p);
        g_bytes -= it->second;
    } else {
        g_free.emplace(it->second, p);
        g_idle += it->second;
    }
    g_live.erase(it);
}

void pool_trim() {
    for (auto& kv : g_free) { cudaFree(kv.second); g_bytes -= kv.first; }
    g_free.clear();
    g_idle = 0;
    for (auto& kv : g_pfree) cudaFreeHost(kv.second);
    g_pfree.clear();
    g_pidle = 0;
}

void* pinned_alloc(size_t bytes) {
    const size_t need = round_up(bytes);
    auto it = g_pfree.lower_bound(need);
    if (it != g_pfree.end() && (it->first <= 2 * need || it->first <= ((size_t)4 << 20))) {
        void* p = it->second;
        g_plive[p] = it->first;
        g_pidle -= it->first;
        g_pfree.erase(it);
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, need);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(QTN_ENOMEM, "page-locked staging buffer of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    g_plive[p] = need;
    return p;
}

void pinned_free(void* p) {
    if (!p) return;
    auto it = g_plive.find(p);
    if (it == g_plive.end()) return;
    if (g_pidle + it->second > kPinnedMaxIdle) cudaFreeHost(p);
    else { g_pfree.emplace(it->second, p); g_pidle += it->second; }
    g_plive.erase(it);
}

void pool_stats(int64_t out[4]) {
    out[0] = g_mallocs;
    out[1] = g_hits;
    out[2] = (int64_t)g_bytes;
    out[3] = (int64_t)g_live.size();
}

}  // namespace qtn

extern "C" int qtn_pool_stats(int64_t out[4]) {
    if (!out) return qtn::fail(QTN_EINVAL, "qtn_pool_stats: null argument");
    qtn::pool_stats(out);
    return QTN_OK;
}
