// Internal declarations shared by the host planner, the executor and the C ABI.
#pragma once
#include <array>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/qaintensor_cuda.h"

namespace qtn {

// Records the message returned by qtn_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);

// The library keeps one stream, one workspace pool and per-plan graphs: device entry points are not re-entrant.
// ApiGuard makes a concurrent call from a second host thread fail (QTN_EBUSY) instead of racing on that state;
// nested calls on the owning thread (entry points built on other entry points) pass.
struct ApiGuard {
    bool ok;
    ApiGuard();
    ~ApiGuard();
};
#define QTN_API_GUARD()                                                                                              \
    qtn::ApiGuard _qtn_api_guard;                                                                                    \
    if (!_qtn_api_guard.ok)                                                                                          \
        return qtn::fail(QTN_EBUSY, "libqaintensor_cuda is not re-entrant: another host thread is inside a device entry point")

// ---- order.cpp ---------------------------------------------------------------
int order_treewidth(int ntensors, int ncontr, const int32_t* pairs, int32_t* perm_out, int32_t* tw_out);
int graph_treewidth(int nv, int ne, const int32_t* edges, int32_t* tw_out, int32_t* ordering_out);
int order_exhaustive(int nt, const int32_t* ranks, const int32_t* const* labels, int nlabels,
                     const int64_t* legdims, int32_t* seq_out, int32_t* nseq_out, int64_t* cost_out);

// ---- plan.cpp ----------------------------------------------------------------
// A two-level additive offset table: offset(i) = lo[i % L] + hi[i / L], where the
// index i enumerates a group of tensor modes (first mode fastest) and L is the
// product of the leading extents.  Positions are element offsets into `tables`.
struct OffTable {
    int64_t n = 1;    // number of indices
    int64_t L = 1;    // size of the lo table
    int64_t lo = 0;   // position of lo[L] in the plan's table buffer
    int64_t hi = 0;   // position of hi[ceil(n / L)]
    int spec = -1;    // index into Plan::table_specs until materialize_tables() fills lo/hi
};
struct TableSpec {
    std::vector<int64_t> extents, strides;
};

enum StepKind { STEP_GEMM = 0, STEP_PERMUTE = 1, STEP_TRACE = 2 };

struct Node {
    std::vector<int> labels;      // layout order, fastest first
    std::vector<int64_t> dims;
    int64_t numel = 1;
    bool is_input = false;
    int input_index = -1;         // which caller tensor (inputs only)
    bool slice_dep = false;       // value differs between slices
    int64_t offset = 0;           // element offset: inputs -> input buffer, others -> arena
    bool persistent = false;      // lives in the slice-invariant arena region
    // inputs only: sliced labels present on this tensor: (position in slice list, element stride)
    std::vector<std::pair<int, int64_t>> slice_strides;
};

struct Step {
    int kind = STEP_GEMM;
    int a = -1, b = -1, out = -1;
    int64_t M = 1, N = 1, K = 1;
    int n_mlabels = 0;            // GEMM: the first n_mlabels labels of the out node come from a
    bool invariant = false;       // independent of the slice id
    bool final_step = false;      // writes (accumulates into) the caller's output
    OffTable a_row, a_k, b_k, b_col, c_row, c_col;
    bool c_dense = false;         // c offset = m + M * n
    bool a_kmajor = false;        // the lowest-stride mode of A is a contracted one (k runs are contiguous, rows are not)
    bool b_kmajor = false;        // same for B
    int split_k = 1;
    int variant = 0;              // kernel tile configuration
    // permute / trace
    OffTable p_tile_in, p_tile_out, p_rest_in, p_rest_out;  // see kernels.cu
    int64_t p_tile = 1, p_rest = 1, t_len = 1;
    int64_t p_tile_smem = 0;      // position of the smem-slot table
};

struct Plan {
    int dtype = QTN_C128;
    int nt = 0;
    std::vector<Node> nodes;
    std::vector<Step> steps;
    std::vector<int> slice_labels;
    std::vector<int64_t> slice_dims;
    int64_t nslices = 1;
    std::vector<int64_t> out_dims;
    int64_t out_numel = 1;
    int final_node = -1;
    std::vector<TableSpec> table_specs;   // what each OffTable enumerates (host-only plans stop here)
    bool tables_ready = false;
    std::vector<int64_t> tables;          // host copy of all offset tables (built on first device use)
    int64_t input_elems = 0;              // packed input buffer size (elements)
    int64_t arena_elems = 0;              // arena size (elements)
    int64_t max_elems = 1;
    double flops = 0, bytes = 0;          // per slice (all steps)
    int launches_per_slice = 0;
    bool dag = false;                     // intermediates never alias: independent steps may overlap
    // device state (exec.cu)
    void* dev = nullptr;
};

int build_plan(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
               const int32_t* order, int norder, const int32_t* slice_labels, int nslice, int dtype,
               Plan** out);
int choose_slices(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                  const int32_t* order, int norder, int max_log2, int64_t min_slices, int32_t* labels_out,
                  int32_t* nlabels_out);
// EXTENSION: randomised greedy order search (see plan.cpp).  order_out: capacity = #contracted labels.
int order_search(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                 int ntrials, uint64_t seed, int max_log2, int32_t* order_out, int32_t* norder_out, double* cost_out);
// Fills Plan::tables from Plan::table_specs (called once, before the first upload).
int materialize_tables(Plan* p);
// Builds the table buffer entries for a mode group; returns the descriptor.
OffTable make_table(std::vector<int64_t>& tables, const std::vector<int64_t>& extents,
                    const std::vector<int64_t>& strides, int64_t max_lo);

// ---- svd.cu --------------------------------------------------------------------
// One SVD problem of a batch (device pointers; double2 = ComplexF64 as void* to keep this header CUDA-free).
struct SvdJob {
    void* A;       // m0 x n0, column-major, overwritten
    int64_t m0, n0;
    void* U;       // m0 x min(m0, n0)
    double* S;     // min(m0, n0)
    void* Vh;      // min(m0, n0) x n0 (ignored when !need_v)
    bool need_v = true;  // false: sigma and U only (m0 >= n0); callers form S*Vh = U^H A themselves
};
// Runs the batch on the library stream and synchronises it; k_out / disc_out are host arrays.
int svd_batched_device(int batch, const SvdJob* jobs, double er, int64_t maxdim, int64_t* k_out, double* disc_out,
                       int* sweeps_out);

// ---- api.cu / orth.cu ----------------------------------------------------------
// C (+)= op(A) op(B), dense column-major ComplexF64 device matrices, on the library stream.
int zgemm_dense(char opa, char opb, int64_t m, int64_t n, int64_t k, const void* dev_a, int64_t lda, const void* dev_b,
                int64_t ldb, void* dev_c, int64_t ldc, bool accumulate);
// Orthonormalises the columns of the m x n (m >= n) device matrix M (ld = m) in place by blocked CholeskyQR2;
// dev_keep holds an untouched copy of M, dev_c (n x n) receives Q^H M.  *ok = false (M and C are then garbage)
// when M is too ill-conditioned for it: callers fall back to the Jacobi SVD.  Synchronises the library stream.
int orth_cholqr2(void* dev_m, const void* dev_keep, void* dev_c, int64_t m, int64_t n, bool* ok);

// ---- pool.cu -------------------------------------------------------------------
// Grow-only, size-cached device workspace (stream-ordered on the library stream; not re-entrant).
void* pool_alloc(size_t bytes);   // nullptr (and the error message set) when the device is out of memory
void pool_free(void* p);          // back to the cache; never cudaFree
void pool_trim();                 // release every cached block to the driver
void* pinned_alloc(size_t bytes); // page-locked host staging buffer from the same kind of cache
void pinned_free(void* p);
void pool_stats(int64_t out[4]);  // cudaMalloc calls, cache hits, bytes owned, blocks handed out
// RAII block of the pool.
struct PoolBuf {
    void* p = nullptr;
    PoolBuf() = default;
    PoolBuf(const PoolBuf&) = delete;
    PoolBuf& operator=(const PoolBuf&) = delete;
    PoolBuf(PoolBuf&& o) noexcept : p(o.p) { o.p = nullptr; }
    int alloc(size_t bytes) {
        if (p) { pool_free(p); p = nullptr; }
        p = pool_alloc(bytes < 256 ? 256 : bytes);
        return p ? QTN_OK : QTN_ENOMEM;
    }
    ~PoolBuf() { if (p) pool_free(p); }
};

// ---- exec.cu -------------------------------------------------------------------
int device_ready();  // QTN_OK or QTN_ENODEVICE (with message)
int plan_device_init(Plan* p);
void plan_device_free(Plan* p);
int plan_upload(Plan* p, const void* const* host_data);
int plan_execute(Plan* p, int64_t s0, int64_t s1, void* dev_out);
int plan_time_steps(Plan* p, int64_t sid, float* ms);
int plan_result_buffer(Plan* p, void** out);
size_t plan_device_bytes(const Plan* p);   // device memory a plan will hold once uploaded (inputs + arena + tables)
void plan_cache_clear();                  // api.cu: drops the plans qtn_contract keeps for repeated identical networks

}  // namespace qtn
