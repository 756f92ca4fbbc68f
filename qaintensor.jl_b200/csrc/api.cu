// extern "C" surface of libqaintensor_cuda (see include/qaintensor_cuda.h).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <memory>
#include <vector>
#include <cstdlib>

#include "kernels.cuh"
#include "qtn_internal.h"

namespace qtn {
const char* last_error_text();
int device_init(int device);
int device_shutdown();
cudaStream_t stream();
int64_t launch_count(int reset);
void count_launch(int64_t n);
int launch_gemm(GemmArgs& g, int variant, int split_k, cudaStream_t st);
int permutedims_device(const void* in, int rank, const int64_t* dims, const int32_t* perm, void* out);
int bench_dmma_peak(double* tflops_out);
}  // namespace qtn

using namespace qtn;

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(QTN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct qtn_plan {
    Plan* p;
};

// ---- plan cache of the one-shot entry point --------------------------------------------------------------------
// `contract(net)` is called again and again on networks of the same structure (parameter sweeps, the reference's own
// `@benchmark contract($T)` loop): the key is the complete structural description (dtype, ranks, dims, labels, order),
// compared exactly, so a hit is the same plan by construction.  A hit skips the planner and the table build, and
// from its second use the plan replays its captured CUDA graph.  Small plans only (<= 256 MB of device memory), at
// most kPlanCacheMax of them, least-recently-used eviction; QTN_PLAN_CACHE=0 disables it; qtn_shutdown / a device
// switch clears it.
namespace {
struct CachedPlan {
    std::vector<int64_t> key;
    std::unique_ptr<Plan> plan;
    uint64_t stamp = 0;
};
std::vector<CachedPlan> g_plan_cache;
uint64_t g_plan_stamp = 0;
const size_t kPlanCacheMax = 4;
const size_t kPlanCacheMaxBytes = (size_t)256 << 20;

bool plan_cache_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("QTN_PLAN_CACHE"); v = e ? atoi(e) : 1; }
    return v != 0;
}
}  // namespace

namespace qtn {
void plan_cache_clear() {
    for (auto& c : g_plan_cache) plan_device_free(c.plan.get());
    g_plan_cache.clear();
}
}  // namespace qtn

extern "C" {

int qtn_version(void) { return 100; }
const char* qtn_last_error(void) { return last_error_text(); }
int qtn_init(int device) { return device_init(device); }
int qtn_shutdown(void) { return device_shutdown(); }
int qtn_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (count) *count = n;
    return QTN_OK;
}
void* qtn_stream(void) { return (void*)stream(); }
int64_t qtn_launch_count(int reset) { return launch_count(reset); }
int qtn_bench_dmma_peak(double* tflops_out) {
    QTN_API_GUARD();
    if (!tflops_out) return fail(QTN_EINVAL, "null argument");
    return bench_dmma_peak(tflops_out);
}

int qtn_order_treewidth(int32_t ntensors, int32_t ncontr, const int32_t* pairs, int32_t* perm_out, int32_t* tw_out) {
    if (!pairs || !perm_out) return fail(QTN_EINVAL, "qtn_order_treewidth: null argument");
    return order_treewidth(ntensors, ncontr, pairs, perm_out, tw_out);
}
int qtn_graph_treewidth(int32_t nv, int32_t ne, const int32_t* edges, int32_t* tw_out, int32_t* ordering_out) {
    return graph_treewidth(nv, ne, edges, tw_out, ordering_out);
}
int qtn_order_exhaustive(int32_t nt, const int32_t* ranks, const int32_t* const* labels, int32_t nlabels,
                         const int64_t* legdims, int32_t* seq_out, int32_t* nseq_out, int64_t* cost_out) {
    if (!ranks || !labels || !legdims || !seq_out || !nseq_out) return fail(QTN_EINVAL, "qtn_order_exhaustive: null argument");
    return order_exhaustive(nt, ranks, labels, nlabels, legdims, seq_out, nseq_out, cost_out);
}

int qtn_plan_create(int32_t nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                    const int32_t* order, int32_t norder, const int32_t* slice_labels, int32_t nslice_labels,
                    int32_t dtype, qtn_plan** plan_out) {
    if (!ranks || !dims || !labels || !plan_out) return fail(QTN_EINVAL, "qtn_plan_create: null argument");
    Plan* p = nullptr;
    int rc = build_plan(nt, ranks, dims, labels, order, norder, slice_labels, nslice_labels, dtype, &p);
    if (rc) return rc;
    *plan_out = new qtn_plan{p};
    return QTN_OK;
}
int qtn_plan_destroy(qtn_plan* plan) {
    QTN_API_GUARD();
    if (!plan) return QTN_OK;
    plan_device_free(plan->p);
    delete plan->p;
    delete plan;
    return QTN_OK;
}
int qtn_choose_slices(int32_t nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                      const int32_t* order, int32_t norder, int32_t max_log2_elems, int64_t min_slices,
                      int32_t* labels_out, int32_t* nlabels_out) {
    if (!ranks || !dims || !labels || !labels_out || !nlabels_out) return fail(QTN_EINVAL, "qtn_choose_slices: null argument");
    return choose_slices(nt, ranks, dims, labels, order, norder, max_log2_elems, min_slices, labels_out, nlabels_out);
}
int qtn_order_search(int32_t nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                     int32_t ntrials, uint64_t seed, int32_t max_log2_elems, int32_t* order_out, int32_t* norder_out,
                     double cost_out[4]) {
    if (!ranks || !dims || !labels || !order_out || !norder_out) return fail(QTN_EINVAL, "qtn_order_search: null argument");
    return order_search(nt, ranks, dims, labels, ntrials, seed, max_log2_elems, order_out, norder_out, cost_out);
}
int qtn_plan_info(const qtn_plan* plan, int64_t info[8], double cost[2]) {
    if (!plan) return fail(QTN_EINVAL, "null plan");
    const Plan& p = *plan->p;
    int64_t inv = 0;
    for (auto& s : p.steps) inv += s.invariant ? 1 : 0;
    if (info) {
        info[0] = (int64_t)p.steps.size();
        info[1] = p.nslices;
        info[2] = (int64_t)p.out_dims.size();
        info[3] = p.out_numel;
        info[4] = p.max_elems;
        info[5] = inv;
        info[6] = p.arena_elems * 16;
        info[7] = p.launches_per_slice + (p.nslices > 1 ? 1 : 0);
    }
    if (cost) { cost[0] = p.flops; cost[1] = p.bytes; }
    return QTN_OK;
}
int qtn_plan_out_dims(const qtn_plan* plan, int64_t* dims_out) {
    if (!plan) return fail(QTN_EINVAL, "null plan");
    for (size_t i = 0; i < plan->p->out_dims.size(); ++i) dims_out[i] = plan->p->out_dims[i];
    return QTN_OK;
}
int qtn_plan_steps(const qtn_plan* plan, int64_t* mnk, int32_t* flags) {
    if (!plan) return fail(QTN_EINVAL, "null plan");
    const Plan& p = *plan->p;
    for (size_t i = 0; i < p.steps.size(); ++i) {
        if (mnk) { mnk[3 * i] = p.steps[i].M; mnk[3 * i + 1] = p.steps[i].N; mnk[3 * i + 2] = p.steps[i].K; }
        if (flags) flags[i] = (p.steps[i].invariant ? 1 : 0) | (p.steps[i].kind << 1) | (p.steps[i].variant << 4) | (p.steps[i].split_k << 8);
    }
    return QTN_OK;
}
int qtn_plan_upload(qtn_plan* plan, const void* const* host_data) {
    QTN_API_GUARD();
    if (!plan || !host_data) return fail(QTN_EINVAL, "qtn_plan_upload: null argument");
    return plan_upload(plan->p, host_data);
}
int qtn_plan_execute(qtn_plan* plan, int64_t slice_begin, int64_t slice_end, void* dev_out) {
    QTN_API_GUARD();
    if (!plan) return fail(QTN_EINVAL, "null plan");
    return plan_execute(plan->p, slice_begin, slice_end, dev_out);
}

static int exec_host(Plan* p, const void* const* host_data, int64_t s0, int64_t s1, void* host_out, bool allreduce);

int qtn_plan_execute_host(qtn_plan* plan, const void* const* host_data, int64_t slice_begin, int64_t slice_end, void* host_out) {
    QTN_API_GUARD();
    if (!plan || !host_out) return fail(QTN_EINVAL, "qtn_plan_execute_host: null argument");
    return exec_host(plan->p, host_data, slice_begin, slice_end, host_out, false);
}
int qtn_plan_time_steps(qtn_plan* plan, int64_t slice_id, float* ms) {
    QTN_API_GUARD();
    if (!plan || !ms) return fail(QTN_EINVAL, "qtn_plan_time_steps: null argument");
    return plan_time_steps(plan->p, slice_id, ms);
}

int qtn_contract(int32_t nt, const void* const* host_data, const int32_t* ranks, const int64_t* const* dims,
                 const int32_t* const* labels, const int32_t* order, int32_t norder, int32_t dtype, void* host_out,
                 int32_t* out_rank, int64_t* out_dims) {
    QTN_API_GUARD();
    if (!host_data || !host_out) return fail(QTN_EINVAL, "qtn_contract: null argument");
    int rc = device_ready();
    if (rc) return rc;
    auto report = [&](const Plan* p) {
        if (out_rank) *out_rank = (int32_t)p->out_dims.size();
        if (out_dims) for (size_t i = 0; i < p->out_dims.size() && i < 64; ++i) out_dims[i] = p->out_dims[i];
    };
    std::vector<int64_t> key;
    const bool cacheable = plan_cache_enabled() && nt > 0 && ranks && dims && labels;
    if (cacheable) {
        key.push_back(dtype);
        key.push_back(nt);
        key.push_back(order ? norder : -1);
        for (int i = 0; i < nt; ++i) {
            key.push_back(ranks[i]);
            if (ranks[i] < 0 || ranks[i] > 60 || (ranks[i] > 0 && (!dims[i] || !labels[i]))) { key.clear(); break; }  // the planner reports it
            for (int j = 0; j < ranks[i]; ++j) { key.push_back(dims[i][j]); key.push_back(labels[i][j]); }
        }
        if (!key.empty() && order && norder > 0) for (int i = 0; i < norder; ++i) key.push_back(order[i]);
        for (auto& c : g_plan_cache)
            if (!key.empty() && c.key == key) {
                c.stamp = ++g_plan_stamp;
                rc = exec_host(c.plan.get(), host_data, 0, 1, host_out, false);
                if (!rc) report(c.plan.get());
                return rc;
            }
    }
    Plan* p = nullptr;
    rc = build_plan(nt, ranks, dims, labels, order, norder, nullptr, 0, dtype, &p);
    if (rc) return rc;
    std::unique_ptr<Plan> holder(p);
    rc = exec_host(p, host_data, 0, 1, host_out, false);
    if (!rc) report(p);
    if (!rc && cacheable && !key.empty() && plan_device_bytes(p) <= kPlanCacheMaxBytes) {
        if (g_plan_cache.size() >= kPlanCacheMax) {
            size_t lru = 0;
            for (size_t i = 1; i < g_plan_cache.size(); ++i) if (g_plan_cache[i].stamp < g_plan_cache[lru].stamp) lru = i;
            plan_device_free(g_plan_cache[lru].plan.get());
            g_plan_cache.erase(g_plan_cache.begin() + (long)lru);
        }
        CachedPlan c;
        c.key = std::move(key);
        c.plan = std::move(holder);
        c.stamp = ++g_plan_stamp;
        g_plan_cache.push_back(std::move(c));
        return rc;
    }
    plan_device_free(p);
    return rc;
}

// ---- NCCL through dlopen (no link-time dependency) -----------------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
static void* g_nccl = nullptr;
static nccl_comm g_comm = nullptr;
static int (*p_ncclGetUniqueId)(nccl_uid*) = nullptr;
static int (*p_ncclCommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
static int (*p_ncclAllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
static const char* (*p_ncclGetErrorString)(int) = nullptr;

static int nccl_load() {
    if (g_nccl) return QTN_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { g_nccl = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl) break; }
    if (!g_nccl) return fail(QTN_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    p_ncclGetUniqueId = (int (*)(nccl_uid*))dlsym(g_nccl, "ncclGetUniqueId");
    p_ncclCommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))dlsym(g_nccl, "ncclCommInitRank");
    p_ncclAllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(g_nccl, "ncclAllReduce");
    p_ncclGetErrorString = (const char* (*)(int))dlsym(g_nccl, "ncclGetErrorString");
    if (!p_ncclGetUniqueId || !p_ncclCommInitRank || !p_ncclAllReduce) return fail(QTN_ENCCL, "libnccl lacks required symbols");
    return QTN_OK;
}
int qtn_nccl_unique_id(void* id_out) {
    int rc = nccl_load();
    if (rc) return rc;
    nccl_uid id;
    int e = p_ncclGetUniqueId(&id);
    if (e) return fail(QTN_ENCCL, "ncclGetUniqueId: %s", p_ncclGetErrorString ? p_ncclGetErrorString(e) : "?");
    memcpy(id_out, &id, 128);
    return QTN_OK;
}
int qtn_nccl_init(int32_t rank, int32_t nranks, const void* id) {
    int rc = nccl_load();
    if (rc) return rc;
    rc = device_ready();
    if (rc) return rc;
    nccl_uid uid;
    memcpy(&uid, id, 128);
    int e = p_ncclCommInitRank(&g_comm, nranks, uid, rank);
    if (e) return fail(QTN_ENCCL, "ncclCommInitRank: %s", p_ncclGetErrorString ? p_ncclGetErrorString(e) : "?");
    return QTN_OK;
}
static int nccl_allreduce(void* dev_buf, int64_t count, int nccl_dtype) {
    if (!g_comm) return fail(QTN_ENCCL, "qtn_nccl_init has not been called");
    int e = p_ncclAllReduce(dev_buf, dev_buf, (size_t)count, nccl_dtype, /*ncclSum*/ 0, g_comm, stream());
    if (e) return fail(QTN_ENCCL, "ncclAllReduce: %s", p_ncclGetErrorString ? p_ncclGetErrorString(e) : "?");
    return QTN_OK;
}
int qtn_nccl_allreduce_sum_f64(void* dev_buf, int64_t count) { return nccl_allreduce(dev_buf, count, /*ncclFloat64*/ 8); }
int qtn_nccl_allreduce_sum_f32(void* dev_buf, int64_t count) { return nccl_allreduce(dev_buf, count, /*ncclFloat32*/ 7); }

static int exec_host(Plan* p, const void* const* host_data, int64_t s0, int64_t s1, void* host_out, bool allreduce) {
    int rc = QTN_OK;
    if (host_data) { rc = plan_upload(p, host_data); if (rc) return rc; }
    if (!p->dev) return fail(QTN_EINVAL, "plan has no uploaded tensors");
    void* out = nullptr;
    if ((rc = plan_result_buffer(p, &out))) return rc;
    rc = plan_execute(p, s0, s1, out);
    const size_t es = p->dtype == QTN_C64 ? 8 : 16;
    if (!rc && allreduce) rc = p->dtype == QTN_C64 ? nccl_allreduce(out, 2 * p->out_numel, /*ncclFloat32*/ 7) : qtn_nccl_allreduce_sum_f64(out, 2 * p->out_numel);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(host_out, out, (size_t)p->out_numel * es, cudaMemcpyDeviceToHost, stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream());
        if (e != cudaSuccess) rc = fail(QTN_ECUDA, "result download failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

int qtn_contract_sliced_range(qtn_plan* plan, const void* const* host_data, int64_t first_slice, int64_t nslices,
                              int32_t rank, int32_t nranks, void* host_out) {
    QTN_API_GUARD();
    if (!plan || !host_out) return fail(QTN_EINVAL, "qtn_contract_sliced: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(QTN_EINVAL, "qtn_contract_sliced: bad rank %d of %d", rank, nranks);
    if (first_slice < 0 || nslices < 0 || first_slice + nslices > plan->p->nslices)
        return fail(QTN_EINVAL, "qtn_contract_sliced: slice range outside the plan's %lld slices", (long long)plan->p->nslices);
    int64_t s0 = first_slice + nslices * rank / nranks, s1 = first_slice + nslices * (rank + 1) / nranks;  // contiguous blocks
    return exec_host(plan->p, host_data, s0, s1, host_out, nranks > 1);
}
int qtn_contract_sliced(qtn_plan* plan, const void* const* host_data, int32_t rank, int32_t nranks, void* host_out) {
    QTN_API_GUARD();
    if (!plan) return fail(QTN_EINVAL, "qtn_contract_sliced: null argument");
    return qtn_contract_sliced_range(plan, host_data, 0, plan->p->nslices, rank, nranks, host_out);
}

// ---- permutedims ---------------------------------------------------------------------
int qtn_permutedims_device(const void* dev_in, int32_t rank, const int64_t* dims, const int32_t* perm, int32_t dtype,
                           void* dev_out) {
    QTN_API_GUARD();
    if (dtype != QTN_C128) return fail(QTN_EINVAL, "only QTN_C128 is implemented");
    int rc = device_ready();
    if (rc) return rc;
    return permutedims_device(dev_in, rank, dims, perm, dev_out);
}
int qtn_permutedims(const void* host_in, int32_t rank, const int64_t* dims, const int32_t* perm, int32_t dtype, void* host_out) {
    QTN_API_GUARD();
    if (dtype != QTN_C128) return fail(QTN_EINVAL, "only QTN_C128 is implemented");
    int rc = device_ready();
    if (rc) return rc;
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
    PoolBuf bin, bout;   // workspace pool (pool.cu)
    if (bin.alloc((size_t)n * 16 + 256) || bout.alloc((size_t)n * 16 + 256)) return QTN_ENOMEM;
    double2 *din = (double2*)bin.p, *dout = (double2*)bout.p;
    cudaMemcpyAsync(din, host_in, (size_t)n * 16, cudaMemcpyHostToDevice, stream());
    rc = permutedims_device(din, rank, dims, perm, dout);
    if (!rc) {
        cudaMemcpyAsync(host_out, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, stream());
        cudaError_t e = cudaStreamSynchronize(stream());
        if (e != cudaSuccess) rc = fail(QTN_ECUDA, "qtn_permutedims: %s", cudaGetErrorString(e));
    }
    return rc;
}

// ---- dense ZGEMM ---------------------------------------------------------------------
}  // extern "C"

namespace qtn {
// C (+)= op(A) op(B) on the library stream with the contraction GEMM kernel in linear-stride mode.
int zgemm_dense(char opa, char opb, int64_t m, int64_t n, int64_t k, const void* dev_a, int64_t lda, const void* dev_b,
                int64_t ldb, void* dev_c, int64_t ldc, bool accumulate) {
    auto ok = [](char c) { return c == 'N' || c == 'T' || c == 'C'; };
    if (!ok(opa) || !ok(opb)) return fail(QTN_EINVAL, "zgemm: op must be N, T or C");
    if (m <= 0 || n <= 0) return QTN_OK;
    if (k <= 0) {
        if (!accumulate) CUDA_TRY(cudaMemset2DAsync(dev_c, (size_t)ldc * 16, 0, (size_t)m * 16, (size_t)n, stream()));
        return QTN_OK;
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = dev_a;
    g.B = dev_b;
    g.C = dev_c;
    auto lin = [](int64_t stride) { TabArg t; t.lo = nullptr; t.hi = nullptr; t.L = stride; t.shift = -1; return t; };
    g.a_row = lin(opa == 'N' ? 1 : lda);
    g.a_k = lin(opa == 'N' ? lda : 1);
    g.b_k = lin(opb == 'N' ? 1 : ldb);
    g.b_col = lin(opb == 'N' ? ldb : 1);
    g.c_row = lin(1);
    g.c_col = lin(ldc);
    g.c_dense = (ldc == m) ? 1 : 0;
    g.conj_a = opa == 'C';
    g.conj_b = opb == 'C';
    g.M = m; g.N = n; g.K = k;
    g.mode = accumulate ? 1 : 0;
    static int pdl = -1;   // QTN_GEMM_PDL=0: plain launches (A/B runs)
    if (pdl < 0) { const char* e = getenv("QTN_GEMM_PDL"); pdl = e ? atoi(e) : 1; }
    g.pdl = pdl;
    return launch_gemm(g, n <= 16 ? 1 : 0, 1, stream());
}
}  // namespace qtn

extern "C" {
int qtn_zgemm_device(char opa, char opb, int64_t m, int64_t n, int64_t k, const void* dev_a, int64_t lda,
                     const void* dev_b, int64_t ldb, void* dev_c, int64_t ldc) {
    QTN_API_GUARD();
    int rc = device_ready();
    if (rc) return rc;
    return zgemm_dense(opa, opb, m, n, k, dev_a, lda, dev_b, ldb, dev_c, ldc, false);
}

}  // extern "C"
