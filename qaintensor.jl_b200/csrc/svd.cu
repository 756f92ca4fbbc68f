// Truncated SVD / MPS entry points (filled in below as the SVD path lands).
#include "qtn_internal.h"
using namespace qtn;
extern "C" {
int qtn_svd_trunc(const void*, int64_t, int64_t, double, int64_t, void*, double*, void*, int64_t*) { return fail(QTN_EINVAL, "qtn_svd_trunc: not implemented yet"); }
int qtn_svd_trunc_batched(int32_t, const void* const*, const int64_t*, const int64_t*, double, int64_t, void* const*, double* const*, void* const*, int64_t*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_svd_trunc_device(void*, int64_t, int64_t, double, int64_t, void*, double*, void*, int64_t*, int32_t*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_contract_svd(const void*, int32_t, const int64_t*, int32_t, const void*, int32_t, const int64_t*, int32_t, double, void*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_mps_create(int32_t, const void* const*, const int64_t*, const int64_t*, int64_t, qtn_mps**) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_mps_destroy(qtn_mps*) { return QTN_OK; }
int qtn_mps_bonds(const qtn_mps*, int64_t*, int64_t*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_mps_download(const qtn_mps*, void* const*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_mps_apply_gate2(qtn_mps*, int32_t, const void*, double, int64_t, double*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_mps_apply_layer(qtn_mps*, int32_t, const int32_t*, const void*, double, int64_t, double*) { return fail(QTN_EINVAL, "not implemented yet"); }
int qtn_mps_overlap(const qtn_mps*, const qtn_mps*, double*) { return fail(QTN_EINVAL, "not implemented yet"); }
}
