/*
 * libqaintensor_cuda -- C ABI of the B200-native contraction / truncation engine
 * that replaces the arithmetic of Qaintensor.jl's hot path.
 *
 * The reference is pure Julia and has no FFI of its own; the seams this ABI cuts
 * at are the Julia call sites listed beside each entry point (paths relative to
 * the reference checkout).  INTEGRATION.md shows the `ccall` stubs.
 *
 * Conventions
 *   - column-major dense arrays; ComplexF64 = two interleaved doubles
 *     (binary-identical to Julia's Complex{Float64} / cuDoubleComplex);
 *   - tensor / leg / contraction indices crossing the ABI are 1-based where the
 *     reference's are (labels follow `contract_rep`, src/contract.jl:39-60:
 *     +k for contraction k, -(i + ncontractions) for open leg i);
 *   - the caller owns every host buffer; the library owns device memory and the
 *     opaque handles; nothing is freed across the boundary;
 *   - every function returns 0 on success, a negative QTN_E* code otherwise;
 *     qtn_last_error() returns the message of the last failure on this thread
 *     (the reference's own error strings where one exists, e.g.
 *     "Error must be positive", src/svd.jl:9);
 *   - there is NO CPU fallback: compute entry points fail with QTN_ENODEVICE
 *     when no sm_100 device is usable.  Host-only entry points (ordering,
 *     planning, cost queries) work without a GPU;
 *   - not re-entrant: call from one host thread at a time.  A second thread entering a device entry point
 *     while another one is inside fails with QTN_EBUSY instead of corrupting the shared stream / workspace state.
 */
#ifndef QAINTENSOR_CUDA_H
#define QAINTENSOR_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTN_OK 0
#define QTN_EINVAL (-1)    /* malformed arguments / network            */
#define QTN_ENODEVICE (-2) /* no usable CUDA device (no CPU fallback)  */
#define QTN_ECUDA (-3)     /* CUDA runtime / kernel failure            */
#define QTN_ENOMEM (-4)    /* device arena exhausted                   */
#define QTN_ENCCL (-5)     /* NCCL unavailable or failed               */
#define QTN_EDOMAIN (-6)   /* reference error() condition (see message) */
#define QTN_EBUSY (-7)     /* another host thread is inside a device entry point (not re-entrant) */

typedef struct qtn_plan qtn_plan;       /* contraction plan (opaque)          */
typedef struct qtn_mps qtn_mps;         /* device-resident MPS (opaque)       */

/* ---- lifetime / diagnostics ------------------------------------------------ */
int qtn_version(void);                           /* 10000*major + 100*minor + patch */
const char* qtn_last_error(void);
/* Selects the device (cudaSetDevice) and creates the library stream.          */
int qtn_init(int device);
int qtn_shutdown(void);
int qtn_device_count(int* count);
/* The stream every kernel of this library is launched on (cudaStream_t).      */
void* qtn_stream(void);
/* Number of this library's kernels launched since the last reset.             */
int64_t qtn_launch_count(int reset);
/* Diagnostics: FP64 tensor-pipe ceiling (TFLOP/s) from a register-only DMMA issue loop;
 * the denominator of the GEMM roofline (MEASURED_PEAKS.json carries no FP64 figure).   */
int qtn_bench_dmma_peak(double* tflops_out);

/* ---- contraction order (host, integer-only, bit-exact) ----------------------
 * Replaces `contraction_order(net)` / `optimize_contraction_order!(net)`
 * (src/network2graph.jl:429-446, 473-479).  pairs[k] = (t1, l1, t2, l2) of
 * contraction k (1-based).  perm_out[ncontr] receives the permutation such that
 * net.contractions = net.contractions[perm].  tw_out (optional) = width of the
 * tree decomposition of the line graph.                                         */
int qtn_order_treewidth(int32_t ntensors, int32_t ncontr, const int32_t* pairs,
                        int32_t* perm_out, int32_t* tw_out);
/* Treewidth heuristic on a plain graph (src/network2graph.jl:300-337), for the
 * reference's known-answer tests: edges[2*ne] 1-based.                          */
int qtn_graph_treewidth(int32_t nv, int32_t ne, const int32_t* edges, int32_t* tw_out,
                        int32_t* ordering_out /* nv, min-fill order, may be NULL */);
/* Replaces `contract_order` (src/contract.jl:184-235): exhaustive cost-capped
 * search.  labels as in qtn_plan_create; legdims[nlabels] = extent of label
 * abs(l) = 1..nlabels.  seq_out holds up to ncontr labels, *nseq_out its length. */
int qtn_order_exhaustive(int32_t nt, const int32_t* ranks, const int32_t* const* labels,
                         int32_t nlabels, const int64_t* legdims, int32_t* seq_out,
                         int32_t* nseq_out, int64_t* cost_out);

/* ---- network builder (host) ----------------------------------------------------
 * The symbolic part of the reference, so that a C / C++ / Julia harness can go from
 * gate matrices to amplitudes through this library alone.  Tensor data
 * is copied into the handle (ComplexF64, column-major); indices are 1-based.
 * qtn_net_create   GeneralTensorNetwork(tensors, contractions, openidx)
 *                  (src/tensor_network.jl:26-33): pairs[k] = (t1, l1, t2, l2),
 *                  openidx[i] = (tensor, leg).
 * qtn_net_tensor_circuit  tensor_circuit!(psi, cgc), non-decomposed branch
 *                  (src/tensor_circuit.jl:44-51): gate g acts on nwires[g] wires
 *                  (concatenated in `wires`), matrices[g] is 2^M x 2^M column-major.
 *                  Error strings of the reference's wire checks are kept.
 * qtn_net_apply_mpo  apply_MPO(psi, mpo::MPO, iwire) (src/mpo.jl:232-252): `op` is an
 *                  operator network whose 2M open legs follow the MPO convention of
 *                  src/mpo.jl:88; returns a NEW network (psi, op untouched).
 * qtn_net_extend_mpo  extend_MPO(mpo::MPO, iwire) (src/mpo.jl:122-157): identity pipes on
 *                  the wires between iwire[M] and iwire[1] (sorted descending); mutates `mpo`.
 * qtn_net_close    EXTENSION: contracts every open leg w with the basis bra <bits[w]|.
 * qtn_net_optimize_order  optimize_contraction_order!(net) (src/network2graph.jl:473-479):
 *                  method 0 = reference treewidth heuristic (bit-exact), 1 = EXTENSION
 *                  qtn_order_search(ntrials, seed, max_log2_elems).
 * qtn_net_contract contract(net) (src/contract.jl:242-264, default-order branch) on the
 *                  GPU; max_log2_elems >= 0 slices the contraction (EXTENSION);
 *                  dtype QTN_C64 rounds the tensors to ComplexF32 (host_out = float pairs).
 * qtn_net_sizes / qtn_net_structure / qtn_net_tensor read the network back
 * (sizes = #tensors, #contractions, #open legs; data pointers stay library-owned).   */
typedef struct qtn_net qtn_net;
int qtn_net_create(int32_t nt, const void* const* host_data, const int32_t* ranks,
                   const int64_t* const* dims, int32_t ncontr, const int32_t* pairs,
                   int32_t nopen, const int32_t* openidx, qtn_net** net_out);
int qtn_net_destroy(qtn_net* net);
int qtn_net_sizes(const qtn_net* net, int32_t sizes[3]);
int qtn_net_structure(const qtn_net* net, int32_t* pairs_out, int32_t* openidx_out);
int qtn_net_tensor(const qtn_net* net, int32_t i, int32_t* rank_out, int64_t* dims_out,
                   const void** data_out);
int qtn_net_tensor_circuit(qtn_net* net, int32_t ngates, const int32_t* nwires,
                           const int32_t* wires, const void* const* matrices);
int qtn_net_apply_mpo(const qtn_net* psi, const qtn_net* op, int32_t nw, const int32_t* iwire,
                      qtn_net** net_out);
int qtn_net_extend_mpo(qtn_net* mpo, int32_t nw, const int32_t* iwire);
int qtn_net_close(qtn_net* net, const int32_t* bits);
int qtn_net_optimize_order(qtn_net* net, int32_t method, int32_t ntrials, uint64_t seed,
                           int32_t max_log2_elems);
int qtn_net_contract(const qtn_net* net, int32_t dtype, int32_t max_log2_elems, void* host_out,
                     int32_t* out_rank, int64_t* out_dims /* capacity 64 */);

/* ---- contraction plans -------------------------------------------------------
 * Replaces `TensorOperations.ncon(tensors, indexlist; order)` as called from
 * src/contract.jl:257, 263.  A plan fixes shapes, labels and order; it can be
 * executed many times with new tensor data of the same shapes.
 *   order == NULL  -> ascending positive labels (ncon default).
 *   slice_labels   -> EXTENSION (no reference counterpart): labels fixed per slice.
 *   dtype          -> QTN_C128 (ComplexF64, FP64 tensor pipe) or QTN_C64 (optional
 *                     ComplexF32 mode on the FP32 pipes; buffers are then float pairs). */
#define QTN_C128 0
#define QTN_C64 1
int qtn_plan_create(int32_t nt, const int32_t* ranks, const int64_t* const* dims,
                    const int32_t* const* labels, const int32_t* order, int32_t norder,
                    const int32_t* slice_labels, int32_t nslice_labels, int32_t dtype,
                    qtn_plan** plan_out);
int qtn_plan_destroy(qtn_plan* plan);
/* Deterministic greedy slice-label choice (rule in DESIGN.md / oracle/plan.py):
 * slice until the largest tensor has <= 2^max_log2_elems elements and there are
 * at least min_slices slices.  Host-only.  labels_out capacity = ncontr.
 * QTN_EDOMAIN (labels found so far still returned) when the target cannot be met:
 * the largest tensor has only open / extent-1 labels left, or too few slices exist. */
int qtn_choose_slices(int32_t nt, const int32_t* ranks, const int64_t* const* dims,
                      const int32_t* const* labels, const int32_t* order, int32_t norder,
                      int32_t max_log2_elems, int64_t min_slices, int32_t* labels_out,
                      int32_t* nlabels_out);
/* EXTENSION (no reference counterpart; the reference's treewidth
 * order of src/network2graph.jl:473-479 stays the default): randomised greedy
 * search for a cheaper pairwise order.  Runs ntrials greedy constructions
 * (deterministic for a given seed), ranks them by flops with a penalty for
 * tensors above 2^max_log2_elems elements (max_log2_elems < 0: no memory target),
 * re-costs the best few exactly with the planner's walk + qtn_choose_slices rule
 * and returns the cheapest as a complete label sequence for `order` of
 * qtn_plan_create / qtn_contract (capacity of order_out = #contracted labels).
 * cost_out (optional): [0] total flops over all slices, [1] flops per slice,
 * [2] number of slices, [3] log2 of the largest tensor per slice.  Host-only.    */
int qtn_order_search(int32_t nt, const int32_t* ranks, const int64_t* const* dims,
                     const int32_t* const* labels, int32_t ntrials, uint64_t seed,
                     int32_t max_log2_elems, int32_t* order_out, int32_t* norder_out,
                     double cost_out[4]);
/* Plan facts (host-only): info[0]=#pairwise steps, [1]=#slices, [2]=output rank,
 * [3]=#output elements, [4]=max tensor elements (per slice), [5]=#slice-invariant
 * steps, [6]=arena bytes, [7]=#kernel launches per slice.
 * cost[0]=sum 8MNK per slice, cost[1]=sum 16(MK+KN+MN) per slice.               */
int qtn_plan_info(const qtn_plan* plan, int64_t info[8], double cost[2]);
int qtn_plan_out_dims(const qtn_plan* plan, int64_t* dims_out /* rank entries */);
/* Per-step shapes: mnk[3*nsteps], flags[nsteps]: bit0 = slice-invariant,
 * bits1-3 = kind (0 pairwise GEMM, 1 permute, 2 partial trace), bits4-7 = kernel
 * tile variant, bits8+ = split-K factor.                                         */
int qtn_plan_steps(const qtn_plan* plan, int64_t* mnk, int32_t* flags);

/* Copy the nt input tensors host -> device (one staged transfer).               */
int qtn_plan_upload(qtn_plan* plan, const void* const* host_data);
/* Run slices [slice_begin, slice_end) on the library stream and ADD their sum to
 * the device buffer dev_out (#output elements, caller-zeroed, device pointer).
 * Asynchronous: returns after enqueueing.                                       */
int qtn_plan_execute(qtn_plan* plan, int64_t slice_begin, int64_t slice_end, void* dev_out);
/* Host-buffer convenience: upload, run slices, download (sum over the slices).  */
int qtn_plan_execute_host(qtn_plan* plan, const void* const* host_data, int64_t slice_begin,
                          int64_t slice_end, void* host_out);
/* Per-step device timing of one slice (CUDA events, ms[nsteps]); diagnostics.   */
int qtn_plan_time_steps(qtn_plan* plan, int64_t slice_id, float* ms);

/* One-shot `ncon`: plan + upload + execute + download (src/contract.jl:257, 263).
 * The library keeps the plans of the last four network structures it was called with
 * (exact match of dtype, ranks, dims, labels and order; <= 256 MB of device memory each):
 * calling it again on the same structure with new tensor data skips the planner and, from
 * the second repeat on, replays the plan's CUDA graph.  QTN_PLAN_CACHE=0 in the environment
 * disables the cache; qtn_shutdown releases it.                                            */
int qtn_contract(int32_t nt, const void* const* host_data, const int32_t* ranks,
                 const int64_t* const* dims, const int32_t* const* labels,
                 const int32_t* order, int32_t norder, int32_t dtype, void* host_out,
                 int32_t* out_rank, int64_t* out_dims /* capacity 64 */);

/* Slice-parallel contraction across ranks: rank r of nranks runs its contiguous
 * block of slices and the partial results are summed with ONE ncclAllReduce.
 * EXTENSION.  Requires qtn_nccl_init.  host_out receives the full sum on all ranks. */
int qtn_nccl_unique_id(void* id_out /* 128 bytes */);
int qtn_nccl_init(int32_t rank, int32_t nranks, const void* id /* 128 bytes */);
int qtn_nccl_allreduce_sum_f64(void* dev_buf, int64_t count);
int qtn_nccl_allreduce_sum_f32(void* dev_buf, int64_t count); /* ComplexF32 mode */
int qtn_contract_sliced(qtn_plan* plan, const void* const* host_data, int32_t rank,
                        int32_t nranks, void* host_out);
/* Same over the slice window [first_slice, first_slice + nslices) only (partial sums,
 * e.g. one batch of a long-running amplitude).                                    */
int qtn_contract_sliced_range(qtn_plan* plan, const void* const* host_data, int64_t first_slice,
                              int64_t nslices, int32_t rank, int32_t nranks, void* host_out);

/* ---- permutedims ---------------------------------------------------------------
 * Replaces Julia `permutedims(A, perm)` at src/contract.jl:244, src/svd.jl:20-21,
 * src/switch.jl:29-35.  out axis i = in axis perm[i] (1-based).                  */
int qtn_permutedims(const void* host_in, int32_t rank, const int64_t* dims,
                    const int32_t* perm, int32_t dtype, void* host_out);
int qtn_permutedims_device(const void* dev_in, int32_t rank, const int64_t* dims,
                           const int32_t* perm, int32_t dtype, void* dev_out);

/* ---- dense ZGEMM (device pointers, column-major) --------------------------------
 * C[m x n] = op(A) * op(B); op = 'N', 'T' or 'C' (adjoint).  Replaces the
 * `*` chains at src/svd.jl:35 and `diagm(S)*adjoint(V)` products.                */
int qtn_zgemm_device(char opa, char opb, int64_t m, int64_t n, int64_t k, const void* dev_a,
                     int64_t lda, const void* dev_b, int64_t ldb, void* dev_c, int64_t ldc);

/* ---- truncated SVD ---------------------------------------------------------------
 * Replaces `LinearAlgebra.svd` + the tail-norm rule of src/svd.jl:29-33
 * (k = n - r* + 1, r* = first r with sqrt(S[n]^2+...+S[n-r+1]^2) > er, strict),
 * then EXTENSION k <- min(k, maxdim) (maxdim <= 0: no cap).  A is m x n column-major
 * (host).  U: m x min(m,n), S: min(m,n), Vh: min(m,n) x n are fully written; the
 * first *k_out columns / values / rows are the kept ones.  er < 0: all values
 * kept (plain `svd`).  If no tail exceeds er, *k_out = 0 (the reference throws).  */
int qtn_svd_trunc(const void* host_a, int64_t m, int64_t n, double er, int64_t maxdim,
                  void* host_u, double* host_s, void* host_vh, int64_t* k_out);
/* Batch of independent problems (ragged shapes allowed).                          */
int qtn_svd_trunc_batched(int32_t batch, const void* const* host_a, const int64_t* m,
                          const int64_t* n, double er, int64_t maxdim, void* const* host_u,
                          double* const* host_s, void* const* host_vh, int64_t* k_out);
/* Device-resident variant: sweeps / final off-diagonal measure reported.          */
int qtn_svd_trunc_device(void* dev_a /* overwritten */, int64_t m, int64_t n, double er,
                         int64_t maxdim, void* dev_u, double* dev_s, void* dev_vh,
                         int64_t* k_out, int32_t* sweeps_out);

/* ---- contract_svd ------------------------------------------------------------------
 * Replaces `contract_svd(T1, T2, (i1, i2); er)` (src/svd.jl:7-38) as a whole.
 * out has dims (T1 without leg i1..., T2 without leg i2...).  Errors:
 * "Error must be positive", "Dimensions of contraction legs do not match".        */
int qtn_contract_svd(const void* host_t1, int32_t rank1, const int64_t* dims1, int32_t i1,
                     const void* host_t2, int32_t rank2, const int64_t* dims2, int32_t i2,
                     double er, void* host_out);

/* ---- device-resident MPS (EXTENSION built from src/switch.jl:18-56) ----------------
 * Site tensors have layout (lbond, 2, rbond) (src/mps.jl:99-110).                  */
int qtn_mps_create(int32_t nsites, const void* const* host_sites, const int64_t* lbond,
                   const int64_t* rbond, int64_t maxdim_capacity, qtn_mps** mps_out);
int qtn_mps_destroy(qtn_mps* mps);
int qtn_mps_bonds(const qtn_mps* mps, int64_t* lbond, int64_t* rbond);
int qtn_mps_download(const qtn_mps* mps, void* const* host_sites);
/* theta = T_i * T_{i+1}; gate (4x4 column-major, index = p_i + 2 p_{i+1}) on the
 * physical legs; SVD; truncate (er, maxdim); T_i <- U, T_{i+1} <- S*V'.
 * site is 1-based.  disc_out = 2-norm of the discarded singular values.           */
int qtn_mps_apply_gate2(qtn_mps* mps, int32_t site, const void* host_gate, double er,
                        int64_t maxdim, double* disc_out);
/* All gates of one brickwork half-layer (disjoint bonds) through the batched SVD.  */
int qtn_mps_apply_layer(qtn_mps* mps, int32_t ngates, const int32_t* sites,
                        const void* host_gates /* ngates 4x4 */, double er, int64_t maxdim,
                        double* disc_out /* ngates */);
/* <a|b> by transfer-matrix contraction; result (re, im).                           */
int qtn_mps_overlap(const qtn_mps* a, const qtn_mps* b, double out[2]);
/* Replaces `MPS(psi::Vector{ComplexF64})` (src/mps.jl:55-89): the whole left-to-right chain of
 * thin SVDs (no truncation) runs on the device.  psi has 2^nsites entries (site 1 = fastest bit).
 * host_sites[i] receives site i: first (2, b1), middle (b_{i-1}, 2, b_i), last (b_{n-1}, 2), column-major;
 * bonds_out[i] (nsites-1 entries) = b_{i+1}.  Buffers must hold min(2^i, 2^(n-i)) bonds.          */
int qtn_mps_from_vector(const void* host_psi, int32_t nsites, void* const* host_sites, int64_t* bonds_out);
/* Replaces the arithmetic of `MPO(m::AbstractMatrix)` (src/mpo.jl:27-90): host_m is the
 * 2^M x 2^M operator (column-major, M = nqubits >= 2).  reshape to fill(2, 2M), permutedims
 * (1, M+1, 2, M+2, ...) (src/mpo.jl:45-50), then the left-to-right chain of un-truncated SVDs
 * (src/mpo.jl:53-66) with the running `diagm(S) * V'` kept on the device.  host_sites[i] receives
 * tensor i+1: (2, 2, b_1), (b_i, 2, 2, b_{i+1}), ..., (b_{M-1}, 2, 2), column-major, with
 * b_i = min(4 b_{i-1}, 4^(M-i)); bonds_out[i] (M-1 entries) = b_{i+1}.  The Summation / openidx
 * bookkeeping (src/mpo.jl:56,67,72-82) stays with the caller.                                      */
int qtn_mpo_from_matrix(const void* host_m, int32_t nqubits, void* const* host_sites, int64_t* bonds_out);
/* Replaces the arithmetic of `decompose!(cg)` (src/decompose.jl:6-52), which is the chain of MPO(m)
 * on the gate matrix; nqubits < 2 fails with the reference's message (src/decompose.jl:7).  The
 * (t, c, w) wire bookkeeping stays with the caller.                                                  */
int qtn_decompose(const void* host_m, int32_t nqubits, void* const* host_sites, int64_t* bonds_out);
/* Replaces `contract_svd_mps(tn; er)` (src/mps.jl:190-201): tcontract = T_1, then
 * tcontract = contract_svd(tcontract, T_j, (ndims(tcontract), 1); er) for j = 2..n.  The contracted legs
 * are the last / first ones, so no permute is needed; every T_j is uploaded once and the running
 * tensor never leaves the device.  numel[j] = length(T_j), first[j] / last[j] = size(T_j, 1) /
 * size(T_j, ndims).  host_out receives numel_out = prod of the open extents elements, column-major
 * in the order size(T_1)[1:end-1] ++ size(T_2)[2:end-1] ++ ... ++ size(T_n)[2:end].  er < 0:
 * QTN_EDOMAIN "Error must be positive" (src/mps.jl:192).  The periodic-boundary check
 * (src/mps.jl:195) stays with the caller, which sees the Summations.                                 */
int qtn_contract_svd_fold(int32_t ntensors, const void* const* host_t, const int64_t* numel,
                          const int64_t* first, const int64_t* last, double er, void* host_out,
                          int64_t numel_out);
/* Replaces the arithmetic of `switch!(mps, i)` (src/switch.jl:18-56) in one call: contract_svd of the
 * two neighbours with er = 0 (:26), exchange of the two physical legs (:28-36), svd (:39),
 * T1' = U, T2' = diagm(S) * V' (:41-52).  T1 is (l1, 2, b) -- or (2, b) when l1 == 0 (first tensor
 * of an MPS with two legs); T2 is (b, 2, r2) -- or (b, 2) when r2 == 0.  host_u receives
 * (l1, 2, bond) [(2, bond)], host_v (bond, 2, r2) [(bond, 2)], *bond_out = min(2 max(l1,1), 2 max(r2,1)). */
int qtn_mps_switch_adjacent(const void* host_t1, int64_t l1, int64_t b, const void* host_t2, int64_t r2,
                            void* host_u, void* host_v, int64_t* bond_out);
/* EXTENSION (SURVEY 8a iii/iv): site-tensor MPO with the layout of src/mpo.jl:66,
 * W_i = (bond_in, out, in, bond_out) = (dl[i], 2, 2, dr[i]), dl[0] = dr[n-1] = 1.
 * qtn_mps_apply_mpo: |psi> <- compress(MPO |psi>): site-wise apply (bonds multiply), a
 * left-to-right gauge sweep that orthogonalises (CholeskyQR2 on the GEMM kernel, verified on the
 * device; U-only Jacobi SVD when the site matrix is too ill-conditioned for it), a right-to-left
 * SVD sweep that truncates with (er, maxdim).  disc_out[n-1] (may be NULL) = discarded 2-norm per bond of the second sweep.
 * qtn_mps_expect_mpo: <psi| MPO |psi> by left-environment contraction, result (re, im).  */
int qtn_mps_apply_mpo(qtn_mps* mps, const void* const* host_mpo_sites, const int64_t* dl,
                      const int64_t* dr, double er, int64_t maxdim, double* disc_out);
int qtn_mps_expect_mpo(const qtn_mps* mps, const void* const* host_mpo_sites, const int64_t* dl,
                       const int64_t* dr, double out[2]);
/* EXTENSION: the gauge step of qtn_mps_apply_mpo on its own.  host_q (m x n, column-major) receives an
 * orthonormal basis of the columns of host_a (m >= n >= 1; where the reference would keep U of
 * `svd`, cf. src/mps.jl:63).  method_out (may be NULL): 1 = blocked CholeskyQR2, 2 = Jacobi SVD
 * fallback (n < 128, a failed orthogonality check, or QTN_ORTH=jacobi in the environment).          */
int qtn_orth_columns(const void* host_a, int64_t m, int64_t n, void* host_q, int32_t* method_out);

/* Diagnostics of the grow-only device workspace pool that serves the transient buffers of the SVD /
 * MPS / MPO / permutedims entry points (where the reference lets Julia's GC own the temporaries of
 * `svd` / `permutedims`, src/svd.jl:22-27): out = { cudaMalloc calls so far, cache hits, bytes owned,
 * blocks currently handed out }.  A steady-state caller sees out[0] stop growing.                    */
int qtn_pool_stats(int64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* QAINTENSOR_CUDA_H */
