// contract_svd (src/svd.jl:7-38) and the device-resident MPS update loop built from
// src/switch.jl:18-56 (two-site theta -> SVD -> split U / S*V'); EXTENSIONS are marked.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "qtn_internal.h"

namespace qtn {
cudaStream_t stream();
void count_launch(int64_t n);
int permutedims_device(const void* in, int rank, const int64_t* dims, const int32_t* perm, void* out);
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(QTN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// out[r, c] = in[r, c] * (by_col ? s[c] : s[r]) for an (rows x cols) block; separate leading dimensions
__global__ void scale_copy_kernel(const double2* __restrict__ in, int64_t ldi, double2* __restrict__ out, int64_t ldo,
                                  int64_t rows, int64_t cols, const double* __restrict__ s, int by_col) {
    const int64_t tot = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e % rows, c = e / rows;
        const double f = s ? (by_col ? s[c] : s[r]) : 1.0;
        const double2 v = in[c * ldi + r];
        out[c * ldo + r] = make_double2(v.x * f, v.y * f);
    }
}

// theta[(l,p1),(p2,r)] <- sum G[(p1',p2'),(p1,p2)] theta ; gate index = p1 + 2*p2, column-major 4x4
__global__ void apply_gate2_kernel(double2* __restrict__ theta, int64_t L, int64_t R, const double2* __restrict__ gate) {
    __shared__ double2 g[16];
    if (threadIdx.x < 16) g[threadIdx.x] = gate[threadIdx.x];
    __syncthreads();
    const int64_t ld = 2 * L;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < L * R; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = e % L, r = e / L;
        double2 v[4], o[4];
#pragma unroll
        for (int p2 = 0; p2 < 2; ++p2)
#pragma unroll
            for (int p1 = 0; p1 < 2; ++p1) v[p1 + 2 * p2] = theta[(p2 + 2 * r) * ld + l + L * p1];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double re = 0, im = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double2 m = g[a + 4 * b];
                re += m.x * v[b].x - m.y * v[b].y;
                im += m.x * v[b].y + m.y * v[b].x;
            }
            o[a] = make_double2(re, im);
        }
#pragma unroll
        for (int p2 = 0; p2 < 2; ++p2)
#pragma unroll
            for (int p1 = 0; p1 < 2; ++p1) theta[(p2 + 2 * r) * ld + l + L * p1] = o[p1 + 2 * p2];
    }
}

// B[(l,a), p', (r,b)] = sum_p W[a,p',p,b] A[l,p,r]   (MPO site applied to an MPS site; l and r fastest)
__global__ void mpo_apply_site_kernel(const double2* __restrict__ A, int64_t L, int64_t R, const double2* __restrict__ W,
                                      int64_t Dl, int64_t Dr, double2* __restrict__ B) {
    const int64_t LB = L * Dl, tot = LB * 2 * R * Dr;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t la = e % LB, pp = (e / LB) % 2, rb = e / (2 * LB);
        const int64_t l = la % L, a = la / L, r = rb % R, b = rb / R;
        double re = 0, im = 0;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const double2 w = W[a + Dl * (pp + 2 * (p + 2 * b))];
            const double2 x = A[l + L * (p + 2 * r)];
            re += w.x * x.x - w.y * x.y;
            im += w.x * x.y + w.y * x.x;
        }
        B[e] = make_double2(re, im);
    }
}

// T2[la, p', b, rb] = sum_{a,p} W[a,p',p,b] T1[la, a, p, rb]      (environment update, middle step)
__global__ void env_apply_w_kernel(const double2* __restrict__ T1, int64_t La, int64_t Rb, const double2* __restrict__ W,
                                   int64_t Dl, int64_t Dr, double2* __restrict__ T2) {
    const int64_t tot = La * 2 * Dr * Rb;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t la = e % La, pp = (e / La) % 2, b = (e / (2 * La)) % Dr, rb = e / (2 * La * Dr);
        double re = 0, im = 0;
        for (int64_t a = 0; a < Dl; ++a)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const double2 w = W[a + Dl * (pp + 2 * (p + 2 * b))];
                const double2 x = T1[la + La * (a + Dl * (p + 2 * rb))];
                re += w.x * x.x - w.y * x.y;
                im += w.x * x.y + w.y * x.x;
            }
        T2[e] = make_double2(re, im);
    }
}

static int scale_copy(const double2* in, int64_t ldi, double2* out, int64_t ldo, int64_t rows, int64_t cols, const double* s, int by_col) {
    if (rows * cols == 0) return QTN_OK;
    int blocks = (int)std::min<int64_t>((rows * cols + 255) / 256, 148 * 8);
    scale_copy_kernel<<<blocks, 256, 0, stream()>>>(in, ldi, out, ldo, rows, cols, s, by_col);
    count_launch(1);
    return cudaGetLastError() == cudaSuccess ? QTN_OK : fail(QTN_ECUDA, "scale_copy launch failed");
}

// transient device buffers come from the grow-only workspace pool (pool.cu): no cudaMalloc / cudaFree per call
using DevBuf = PoolBuf;


// M (mm x nn, mm >= nn, device) <- an orthonormal basis Q of its columns; keep <- the original M;
// carry (nn x nn) <- Q^H M, so that Q * carry = M.  A gauge sweep needs no singular values: CholeskyQR2 (GEMM
// work) when the matrix is large enough to pay, the U-only Jacobi SVD when its checks fail or QTN_ORTH=jacobi
// is set (A/B runs).  u_scratch: mm x nn elements, s_scratch: nn doubles.  *method: 1 CholeskyQR2, 2 Jacobi.
static int orth_columns(void* M, void* keep, void* carry, int64_t mm, int64_t nn, void* u_scratch, double* s_scratch, int* method) {
    cudaStream_t st = stream();
    int rc;
    CUDA_TRY(cudaMemcpyAsync(keep, M, (size_t)mm * nn * 16, cudaMemcpyDeviceToDevice, st));
    const char* env = getenv("QTN_ORTH");
    bool done = false;
    if (nn >= 128 && !(env && strcmp(env, "jacobi") == 0)) {
        if ((rc = orth_cholqr2(M, keep, carry, mm, nn, &done))) return rc;
        if (!done) CUDA_TRY(cudaMemcpyAsync(M, keep, (size_t)mm * nn * 16, cudaMemcpyDeviceToDevice, st));
    }
    if (method) *method = done ? 1 : 2;
    if (done) return QTN_OK;
    SvdJob job{M, mm, nn, u_scratch, s_scratch, nullptr};
    job.need_v = false;
    int64_t k = 0;
    if ((rc = svd_batched_device(1, &job, -1.0, 0, &k, nullptr, nullptr))) return rc;
    if ((rc = scale_copy((const double2*)u_scratch, mm, (double2*)M, mm, mm, nn, nullptr, 1))) return rc;
    return zgemm_dense('C', 'N', nn, nn, mm, M, mm, keep, mm, carry, nn, false);
}


// The arithmetic of contract_svd (src/svd.jl:22-35) on device-resident, already permuted operands:
// a1 (m1 x D) and a2 (D x m2), column-major, both OVERWRITTEN (the Jacobi works in place);
// out (m1 x m2) <- U1[:, :k1] diag(S1) V1h[:k1, :] U2[:, :k2] diag(S2) V2h[:k2, :] with the tail-norm cutoff `er`.
// The two SVDs run as one batch.  Work on the library stream; returns after the SVD's host synchronisation,
// the closing GEMMs may still be in flight.
int contract_svd_device(void* a1, int64_t m1, void* a2, int64_t m2, int64_t D, double er, void* out) {
    int rc;
    const int64_t r1 = std::min(m1, D), r2 = std::min(D, m2);
    DevBuf u1, u2, v1, v2, s1, s2, x1, x2, y, z;
    if ((rc = u1.alloc(m1 * r1 * 16)) || (rc = v1.alloc(r1 * D * 16)) || (rc = s1.alloc(r1 * 8)) ||
        (rc = u2.alloc(D * r2 * 16)) || (rc = v2.alloc(r2 * m2 * 16)) || (rc = s2.alloc(r2 * 8)))
        return rc;
    SvdJob jobs[2] = {{(double2*)a1, m1, D, (double2*)u1.p, (double*)s1.p, (double2*)v1.p},
                      {(double2*)a2, D, m2, (double2*)u2.p, (double*)s2.p, (double2*)v2.p}};
    int64_t k[2] = {0, 0};
    if ((rc = svd_batched_device(2, jobs, er, 0, k, nullptr, nullptr))) return rc;
    if (k[0] == 0 || k[1] == 0)
        return fail(QTN_EDOMAIN, "contract_svd: the cutoff removes every singular value (the reference's findfirst returns nothing)");
    const int64_t k1 = k[0], k2 = k[1];
    if ((rc = x1.alloc(m1 * k1 * 16)) || (rc = x2.alloc(k2 * m2 * 16)) || (rc = y.alloc(k1 * k2 * 16))) return rc;
    if ((rc = scale_copy((double2*)u1.p, m1, (double2*)x1.p, m1, m1, k1, (double*)s1.p, 1))) return rc;
    if ((rc = scale_copy((double2*)v2.p, r2, (double2*)x2.p, k2, k2, m2, (double*)s2.p, 0))) return rc;
    if ((rc = qtn_zgemm_device('N', 'N', k1, k2, D, v1.p, r1, u2.p, D, y.p, k1))) return rc;
    if (m1 * k1 * k2 + m1 * k2 * m2 <= k1 * k2 * m2 + m1 * k1 * m2) {
        if ((rc = z.alloc(m1 * k2 * 16))) return rc;
        if ((rc = qtn_zgemm_device('N', 'N', m1, k2, k1, x1.p, m1, y.p, k1, z.p, m1))) return rc;
        if ((rc = qtn_zgemm_device('N', 'N', m1, m2, k2, z.p, m1, x2.p, k2, out, m1))) return rc;
    } else {
        if ((rc = z.alloc(k1 * m2 * 16))) return rc;
        if ((rc = qtn_zgemm_device('N', 'N', k1, m2, k2, y.p, k1, x2.p, k2, z.p, k1))) return rc;
        if ((rc = qtn_zgemm_device('N', 'N', m1, m2, k1, x1.p, m1, z.p, k1, out, m1))) return rc;
    }
    return QTN_OK;
}

// Sequential-SVD chain shared by MPS(psi) (src/mps.jl:55-89, d = 2), MPO(m) (src/mpo.jl:40-75, d = 4) and
// decompose! (src/decompose.jl:17-48, d = 4): rest (d x cols, device, overwritten) is split site by site,
//   rest -> reshape(lbond * d, :) -> svd -> site_i = U, rest = diag(S) V',
// the running `rest` never leaves the device; every site is copied to its host buffer as it is produced.
// tmp: scratch of the size of rest.  bonds_out[nsites - 1].
int svd_chain_device(void* rest, void* tmp, int64_t total, int64_t d, int nsites, void* const* host_sites, int64_t* bonds_out) {
    cudaStream_t st = stream();
    int rc;
    int64_t lbond = 1, cols = total;
    // the factors of the largest step bound every step's: allocate once
    int64_t maxU = 0, maxV = 0, maxS = 0;
    {
        int64_t lb = 1, c = total;
        for (int i = 1; i < nsites; ++i) {
            const int64_t m = lb * d;
            c /= d;
            const int64_t r = std::min(m, c);
            maxU = std::max(maxU, m * r); maxV = std::max(maxV, r * c); maxS = std::max(maxS, r);
            lb = r;
        }
    }
    DevBuf u, s, v;
    if ((rc = u.alloc(maxU * 16)) || (rc = s.alloc(maxS * 8)) || (rc = v.alloc(maxV * 16))) return rc;
    for (int i = 1; i < nsites; ++i) {
        const int64_t m = lbond * d;
        cols /= d;
        const int64_t r = std::min(m, cols);
        SvdJob job{(double2*)rest, m, cols, (double2*)u.p, (double*)s.p, (double2*)v.p};
        int64_t k = 0;
        if ((rc = svd_batched_device(1, &job, -1.0, 0, &k, nullptr, nullptr))) return rc;
        CUDA_TRY(cudaMemcpyAsync(host_sites[i - 1], u.p, (size_t)m * r * 16, cudaMemcpyDeviceToHost, st));
        if ((rc = scale_copy((double2*)v.p, r, (double2*)tmp, r, r, cols, (double*)s.p, 0))) return rc;
        std::swap(rest, tmp);
        bonds_out[i - 1] = r;
        lbond = r;
    }
    CUDA_TRY(cudaMemcpyAsync(host_sites[nsites - 1], rest, (size_t)lbond * d * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return QTN_OK;
}

}  // namespace qtn

using namespace qtn;

struct qtn_mps {
    int n = 0;
    int64_t cap = 0;                  // bond capacity
    std::vector<double2*> site;       // each 2 * cap * cap elements
    std::vector<int64_t> lb, rb;
    double2* theta = nullptr;         // batch scratch: per bond theta (2cap x 2cap), U, Vh, S
    double2* ubuf = nullptr;
    double2* vbuf = nullptr;
    double* sbuf = nullptr;
    double2* gates = nullptr;
    int scratch_bonds = 0;
};

extern "C" {

int qtn_contract_svd(const void* host_t1, int32_t rank1, const int64_t* dims1, int32_t i1, const void* host_t2,
                     int32_t rank2, const int64_t* dims2, int32_t i2, double er, void* host_out) {
    QTN_API_GUARD();
    if (!(er >= 0)) return fail(QTN_EDOMAIN, "Error must be positive");
    if (!host_t1 || !host_t2 || !host_out || !dims1 || !dims2) return fail(QTN_EINVAL, "qtn_contract_svd: null argument");
    if (i1 < 1 || i2 < 1) return fail(QTN_EINVAL, "qtn_contract_svd: leg index must be >= 1");
    const int64_t D1 = i1 <= rank1 ? dims1[i1 - 1] : 1, D2 = i2 <= rank2 ? dims2[i2 - 1] : 1;  // Julia size(A, d > ndims) == 1
    if (D1 != D2) return fail(QTN_EDOMAIN, "Dimensions of contraction legs do not match");
    if (i1 > rank1 || i2 > rank2) return fail(QTN_EINVAL, "qtn_contract_svd: leg index beyond the tensor rank");
    int rc = device_ready();
    if (rc) return rc;
    cudaStream_t st = stream();
    int64_t n1 = 1, n2 = 1;
    for (int i = 0; i < rank1; ++i) n1 *= dims1[i];
    for (int i = 0; i < rank2; ++i) n2 *= dims2[i];
    const int64_t D = D1, m1 = n1 / D, m2 = n2 / D;
    if (n1 == 0 || n2 == 0) return fail(QTN_EINVAL, "qtn_contract_svd: empty tensor");
    DevBuf t1, t2, p1, p2, out;
    if ((rc = t1.alloc(n1 * 16)) || (rc = t2.alloc(n2 * 16)) || (rc = p1.alloc(n1 * 16)) || (rc = p2.alloc(n2 * 16)) ||
        (rc = out.alloc(m1 * m2 * 16)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(t1.p, host_t1, n1 * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t2.p, host_t2, n2 * 16, cudaMemcpyHostToDevice, st));
    // src/svd.jl:20-21: T1 -> (others..., i1), T2 -> (i2, others...)
    std::vector<int32_t> perm1, perm2;
    for (int a = 1; a <= rank1; ++a) if (a != i1) perm1.push_back(a);
    perm1.push_back(i1);
    perm2.push_back(i2);
    for (int a = 1; a <= rank2; ++a) if (a != i2) perm2.push_back(a);
    if ((rc = permutedims_device(t1.p, rank1, dims1, perm1.data(), p1.p))) return rc;
    if ((rc = permutedims_device(t2.p, rank2, dims2, perm2.data(), p2.p))) return rc;
    if ((rc = contract_svd_device(p1.p, m1, p2.p, m2, D, er, out.p))) return rc;
    CUDA_TRY(cudaMemcpyAsync(host_out, out.p, m1 * m2 * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return QTN_OK;
}

// ---------------- device-resident MPS (EXTENSION) -------------------------------------------------
int qtn_mps_create(int32_t nsites, const void* const* host_sites, const int64_t* lbond, const int64_t* rbond,
                   int64_t maxdim_capacity, qtn_mps** mps_out) {
    QTN_API_GUARD();
    if (nsites < 2 || !host_sites || !lbond || !rbond || !mps_out || maxdim_capacity < 1) return fail(QTN_EINVAL, "qtn_mps_create: bad argument");
    int rc = device_ready();
    if (rc) return rc;
    for (int i = 0; i < nsites; ++i) {
        if (lbond[i] < 1 || rbond[i] < 1 || lbond[i] > maxdim_capacity || rbond[i] > maxdim_capacity)
            return fail(QTN_EINVAL, "qtn_mps_create: bond of site %d outside [1, capacity]", i + 1);
        if (i > 0 && lbond[i] != rbond[i - 1]) return fail(QTN_EINVAL, "qtn_mps_create: bonds of sites %d and %d do not match", i, i + 1);
    }
    qtn_mps* m = new qtn_mps();
    m->n = nsites;
    m->cap = maxdim_capacity;
    m->site.assign(nsites, nullptr);
    m->lb.assign(lbond, lbond + nsites);
    m->rb.assign(rbond, rbond + nsites);
    const size_t sb = (size_t)2 * m->cap * m->cap * 16;
    for (int i = 0; i < nsites; ++i) {
        if (!host_sites[i]) { qtn_mps_destroy(m); return fail(QTN_EINVAL, "qtn_mps_create: site %d is a null pointer", i + 1); }
        if (cudaMalloc((void**)&m->site[i], sb) != cudaSuccess) { qtn_mps_destroy(m); return fail(QTN_ENOMEM, "qtn_mps_create: out of device memory"); }
        cudaError_t e = cudaMemcpyAsync(m->site[i], host_sites[i], (size_t)lbond[i] * 2 * rbond[i] * 16, cudaMemcpyHostToDevice, stream());
        if (e != cudaSuccess) { qtn_mps_destroy(m); return fail(QTN_ECUDA, "qtn_mps_create: upload of site %d failed: %s", i + 1, cudaGetErrorString(e)); }
    }
    {
        cudaError_t e = cudaStreamSynchronize(stream());
        if (e != cudaSuccess) { qtn_mps_destroy(m); return fail(QTN_ECUDA, "qtn_mps_create: %s", cudaGetErrorString(e)); }
    }
    *mps_out = m;
    return QTN_OK;
}

int qtn_mps_destroy(qtn_mps* m) {
    QTN_API_GUARD();
    if (!m) return QTN_OK;
    if (stream()) cudaStreamSynchronize(stream());
    for (auto p : m->site) if (p) cudaFree(p);
    if (m->theta) cudaFree(m->theta);
    if (m->ubuf) cudaFree(m->ubuf);
    if (m->vbuf) cudaFree(m->vbuf);
    if (m->sbuf) cudaFree(m->sbuf);
    if (m->gates) cudaFree(m->gates);
    delete m;
    return QTN_OK;
}

int qtn_mps_bonds(const qtn_mps* m, int64_t* lbond, int64_t* rbond) {
    if (!m) return fail(QTN_EINVAL, "null mps");
    for (int i = 0; i < m->n; ++i) { if (lbond) lbond[i] = m->lb[i]; if (rbond) rbond[i] = m->rb[i]; }
    return QTN_OK;
}

int qtn_mps_download(const qtn_mps* m, void* const* host_sites) {
    QTN_API_GUARD();
    if (!m || !host_sites) return fail(QTN_EINVAL, "null argument");
    for (int i = 0; i < m->n; ++i)
        CUDA_TRY(cudaMemcpyAsync(host_sites[i], m->site[i], (size_t)m->lb[i] * 2 * m->rb[i] * 16, cudaMemcpyDeviceToHost, stream()));
    CUDA_TRY(cudaStreamSynchronize(stream()));
    return QTN_OK;
}

static int mps_scratch(qtn_mps* m, int bonds) {
    if (bonds <= m->scratch_bonds) return QTN_OK;
    cudaStreamSynchronize(stream());
    if (m->theta) cudaFree(m->theta);
    if (m->ubuf) cudaFree(m->ubuf);
    if (m->vbuf) cudaFree(m->vbuf);
    if (m->sbuf) cudaFree(m->sbuf);
    if (m->gates) cudaFree(m->gates);
    m->theta = m->ubuf = m->vbuf = nullptr; m->sbuf = nullptr; m->gates = nullptr;
    const size_t mat = (size_t)4 * m->cap * m->cap * 16;
    if (cudaMalloc((void**)&m->theta, mat * bonds) != cudaSuccess || cudaMalloc((void**)&m->ubuf, mat * bonds) != cudaSuccess ||
        cudaMalloc((void**)&m->vbuf, mat * bonds) != cudaSuccess || cudaMalloc((void**)&m->sbuf, (size_t)2 * m->cap * 8 * bonds) != cudaSuccess ||
        cudaMalloc((void**)&m->gates, (size_t)256 * bonds) != cudaSuccess) {
        m->scratch_bonds = 0;
        return fail(QTN_ENOMEM, "qtn_mps: scratch for %d bonds does not fit", bonds);
    }
    m->scratch_bonds = bonds;
    return QTN_OK;
}

int qtn_mps_apply_layer(qtn_mps* m, int32_t ngates, const int32_t* sites, const void* host_gates, double er, int64_t maxdim,
                        double* disc_out) {
    QTN_API_GUARD();
    if (!m || !sites || !host_gates || ngates < 0) return fail(QTN_EINVAL, "qtn_mps_apply_layer: null argument");
    if (ngates == 0) return QTN_OK;
    if (maxdim <= 0 || maxdim > m->cap) maxdim = m->cap;
    std::vector<char> used(m->n, 0);
    for (int g = 0; g < ngates; ++g) {
        const int s = sites[g];
        if (s < 1 || s >= m->n) return fail(QTN_EINVAL, "qtn_mps_apply_layer: gate %d acts on sites (%d, %d) outside 1..%d", g + 1, s, s + 1, m->n);
        if (used[s - 1] || used[s]) return fail(QTN_EINVAL, "qtn_mps_apply_layer: gates of one layer must act on disjoint bonds");
        used[s - 1] = used[s] = 1;
    }
    int rc = mps_scratch(m, ngates);
    if (rc) return rc;
    cudaStream_t st = stream();
    const size_t mat = (size_t)4 * m->cap * m->cap;
    CUDA_TRY(cudaMemcpyAsync(m->gates, host_gates, (size_t)256 * ngates, cudaMemcpyHostToDevice, st));
    std::vector<SvdJob> jobs(ngates);
    for (int g = 0; g < ngates; ++g) {
        const int i = sites[g] - 1;
        const int64_t L = m->lb[i], B = m->rb[i], R = m->rb[i + 1];
        double2* th = m->theta + mat * g;
        // theta = T_i (2L x B) * T_{i+1} (B x 2R)      (src/switch.jl:26 with er = 0, no SVD round trip)
        if ((rc = qtn_zgemm_device('N', 'N', 2 * L, 2 * R, B, m->site[i], 2 * L, m->site[i + 1], B, th, 2 * L))) return rc;
        int blocks = (int)std::min<int64_t>((L * R + 255) / 256, 148 * 4);
        apply_gate2_kernel<<<blocks, 256, 0, st>>>(th, L, R, m->gates + 16 * g);
        count_launch(1);
        jobs[g] = SvdJob{th, 2 * L, 2 * R, m->ubuf + mat * g, m->sbuf + 2 * m->cap * g, m->vbuf + mat * g};
        if (L >= R) {
            // U-only SVD: the right site S*V'[:k] equals U[:, :k]^H theta exactly, so V is never accumulated
            // (halves the rotation updates); theta is saved in the V scratch because the Jacobi overwrites it
            jobs[g].need_v = false;
            CUDA_TRY(cudaMemcpyAsync(m->vbuf + mat * g, th, (size_t)4 * L * R * 16, cudaMemcpyDeviceToDevice, st));
        }
    }
    std::vector<int64_t> k(ngates);
    if ((rc = svd_batched_device(ngates, jobs.data(), er, maxdim, k.data(), disc_out, nullptr))) return rc;
    for (int g = 0; g < ngates; ++g) {
        const int i = sites[g] - 1;
        const int64_t L = m->lb[i], R = m->rb[i + 1], rfull = std::min(2 * L, 2 * R);
        int64_t kk = std::max<int64_t>(k[g], 1);  // keep at least one state
        // T_i <- U[:, :k] (L,2,k);  T_{i+1} <- diag(S) V'[:k, :] (k,2,R)      (src/switch.jl:50-52)
        if ((rc = scale_copy(m->ubuf + mat * g, 2 * L, m->site[i], 2 * L, 2 * L, kk, nullptr, 1))) return rc;
        if (jobs[g].need_v) {
            if ((rc = scale_copy(m->vbuf + mat * g, rfull, m->site[i + 1], kk, kk, 2 * R, m->sbuf + 2 * m->cap * g, 0))) return rc;
        } else {
            if ((rc = qtn_zgemm_device('C', 'N', kk, 2 * R, 2 * L, m->ubuf + mat * g, 2 * L, m->vbuf + mat * g, 2 * L, m->site[i + 1], kk))) return rc;
        }
        m->rb[i] = kk;
        m->lb[i + 1] = kk;
    }
    return QTN_OK;
}

int qtn_mps_apply_gate2(qtn_mps* m, int32_t site, const void* host_gate, double er, int64_t maxdim, double* disc_out) {
    QTN_API_GUARD();
    return qtn_mps_apply_layer(m, 1, &site, host_gate, er, maxdim, disc_out);
}

int qtn_mps_overlap(const qtn_mps* a, const qtn_mps* b, double out[2]) {
    QTN_API_GUARD();
    if (!a || !b || !out) return fail(QTN_EINVAL, "null argument");
    if (a->n != b->n) return fail(QTN_EINVAL, "qtn_mps_overlap: different lengths");
    if (a->lb[0] != 1 || b->lb[0] != 1 || a->rb[a->n - 1] != 1 || b->rb[b->n - 1] != 1)
        return fail(QTN_EINVAL, "qtn_mps_overlap: boundary bonds must be 1");
    int rc = device_ready();
    if (rc) return rc;
    const int64_t cap = std::max(a->cap, b->cap);
    DevBuf E, E2, Y;
    if ((rc = E.alloc(cap * cap * 16)) || (rc = E2.alloc(cap * cap * 16)) || (rc = Y.alloc(2 * cap * cap * 16))) return rc;
    const double2 one = make_double2(1.0, 0.0);
    CUDA_TRY(cudaMemcpyAsync(E.p, &one, 16, cudaMemcpyHostToDevice, stream()));
    void *e = E.p, *e2 = E2.p;
    for (int i = 0; i < a->n; ++i) {
        const int64_t la = a->lb[i], ra = a->rb[i], lb = b->lb[i], rb = b->rb[i];
        // Y (la x 2rb) = E (la x lb) B_i (lb x 2rb);  E' (ra x rb) = A_i^H (ra x 2la) Y (2la x rb)
        if ((rc = qtn_zgemm_device('N', 'N', la, 2 * rb, lb, e, la, b->site[i], lb, Y.p, la))) return rc;
        if ((rc = qtn_zgemm_device('C', 'N', ra, rb, 2 * la, a->site[i], 2 * la, Y.p, 2 * la, e2, ra))) return rc;
        std::swap(e, e2);
    }
    double2 res;
    CUDA_TRY(cudaMemcpyAsync(&res, e, 16, cudaMemcpyDeviceToHost, stream()));
    CUDA_TRY(cudaStreamSynchronize(stream()));
    out[0] = res.x;
    out[1] = res.y;
    return QTN_OK;
}

// ---------------- MPS(psi) on the device (src/mps.jl:55-89) ----------------------------------------------
int qtn_mps_from_vector(const void* host_psi, int32_t nsites, void* const* host_sites, int64_t* bonds_out) {
    QTN_API_GUARD();
    if (!host_psi || !host_sites || !bonds_out) return fail(QTN_EINVAL, "qtn_mps_from_vector: null argument");
    if (nsites < 2 || nsites > 30) return fail(QTN_EINVAL, "qtn_mps_from_vector: nsites must be in 2..30");
    int rc = device_ready();
    if (rc) return rc;
    const int64_t n = (int64_t)1 << nsites;
    DevBuf rest, tmp;
    if ((rc = rest.alloc(n * 16)) || (rc = tmp.alloc(n * 16))) return rc;
    CUDA_TRY(cudaMemcpyAsync(rest.p, host_psi, n * 16, cudaMemcpyHostToDevice, stream()));
    return svd_chain_device(rest.p, tmp.p, n, 2, nsites, host_sites, bonds_out);
}

// ---------------- MPO(m) / decompose! on the device (src/mpo.jl:27-90, src/decompose.jl:6-52) ---------------
// host_m: 2^M x 2^M operator, column-major.  reshape(m, fill(2, 2M)) -> permutedims (1, M+1, 2, M+2, ...) ->
// the sequential-SVD chain with physical dimension 4.  host_sites[i] receives site i+1: (bond_i, 2, 2, bond_{i+1})
// with bond_0 = 1 (the first site is (2, 2, bond_1), the last (bond_{M-1}, 2, 2)); the svd is not truncated, so
// bond_i = min(4 bond_{i-1}, 4^(M-i)) is known to the caller, who sizes the buffers.  bonds_out[M - 1].
int qtn_mpo_from_matrix(const void* host_m, int32_t nqubits, void* const* host_sites, int64_t* bonds_out) {
    QTN_API_GUARD();
    if (!host_m || !host_sites || !bonds_out) return fail(QTN_EINVAL, "qtn_mpo_from_matrix: null argument");
    if (nqubits < 2) return fail(QTN_EDOMAIN, "Need at least two qubits to split (a one-qubit operator is its own MPO)");
    if (nqubits > 15) return fail(QTN_EINVAL, "qtn_mpo_from_matrix: nqubits must be in 2..15");
    int rc = device_ready();
    if (rc) return rc;
    const int M = nqubits;
    const int64_t n = (int64_t)1 << (2 * M);
    DevBuf in, rest;
    if ((rc = in.alloc(n * 16)) || (rc = rest.alloc(n * 16))) return rc;
    CUDA_TRY(cudaMemcpyAsync(in.p, host_m, n * 16, cudaMemcpyHostToDevice, stream()));
    std::vector<int64_t> dims(2 * M, 2);
    std::vector<int32_t> perm;
    for (int i = 1; i <= M; ++i) { perm.push_back(i); perm.push_back(i + M); }   // src/mpo.jl:45-50, src/decompose.jl:19-23
    if ((rc = permutedims_device(in.p, 2 * M, dims.data(), perm.data(), rest.p))) return rc;
    return svd_chain_device(rest.p, in.p, n, 4, M, host_sites, bonds_out);
}

// decompose!(cg) (src/decompose.jl:6-52) performs exactly the chain of MPO(m) on the gate's matrix; the wire /
// bond bookkeeping (t, c, w) stays with the caller.
int qtn_decompose(const void* host_m, int32_t nqubits, void* const* host_sites, int64_t* bonds_out) {
    QTN_API_GUARD();
    if (nqubits < 2) return fail(QTN_EDOMAIN, "Only decompose Circuit Gates that apply to multiple wires");
    return qtn_mpo_from_matrix(host_m, nqubits, host_sites, bonds_out);
}

// contract_svd_mps (src/mps.jl:190-201): tcontract = T_1; tcontract = contract_svd(tcontract, T_j, (ndims, 1); er)
// for j = 2..n.  The contracted legs are already the last / the first, so the fold needs no permute: every T_j is
// uploaded once, the running tensor stays on the device, one download at the end.
// numel[j] = elements of T_j, first[j] / last[j] = its first / last extent.  host_out: numel_out elements
// (= prod of all open extents; the caller knows the shape: dims(T_1)[:-1] ++ dims(T_2)[2:-1] ++ ... ++ dims(T_n)[2:]).
int qtn_contract_svd_fold(int32_t ntensors, const void* const* host_t, const int64_t* numel, const int64_t* first,
                          const int64_t* last, double er, void* host_out, int64_t numel_out) {
    QTN_API_GUARD();
    if (!(er >= 0)) return fail(QTN_EDOMAIN, "Error must be positive");
    if (ntensors < 1 || !host_t || !numel || !first || !last || !host_out) return fail(QTN_EINVAL, "qtn_contract_svd_fold: null argument");
    int64_t tot = 0;     // elements of the running tensor
    int64_t peak = 0;
    for (int j = 0; j < ntensors; ++j) {
        if (!host_t[j] || numel[j] < 1 || first[j] < 1 || last[j] < 1 || numel[j] % first[j] || numel[j] % last[j])
            return fail(QTN_EINVAL, "qtn_contract_svd_fold: tensor %d has an inconsistent shape", j + 1);
        if (j == 0) tot = numel[0];
        else {
            if (first[j] != last[j - 1]) return fail(QTN_EDOMAIN, "Dimensions of contraction legs do not match");
            tot = tot / last[j - 1] * (numel[j] / first[j]);
        }
        peak = std::max(peak, tot);
        if (peak > ((int64_t)1 << 33)) return fail(QTN_EINVAL, "qtn_contract_svd_fold: the running tensor exceeds 2^33 elements");
    }
    if (tot != numel_out) return fail(QTN_EINVAL, "qtn_contract_svd_fold: output has %lld elements, expected %lld", (long long)numel_out, (long long)tot);
    int rc = device_ready();
    if (rc) return rc;
    cudaStream_t st = stream();
    DevBuf acc, nxt, tj;
    int64_t maxt = 0;
    for (int j = 1; j < ntensors; ++j) maxt = std::max(maxt, numel[j]);
    if ((rc = acc.alloc(peak * 16)) || (rc = nxt.alloc(peak * 16)) || (rc = tj.alloc(std::max<int64_t>(maxt, 1) * 16))) return rc;
    CUDA_TRY(cudaMemcpyAsync(acc.p, host_t[0], (size_t)numel[0] * 16, cudaMemcpyHostToDevice, st));
    void *a = acc.p, *b = nxt.p;
    int64_t cur = numel[0];   // elements of the running tensor
    for (int j = 1; j < ntensors; ++j) {
        // running tensor as (m1 x D), T_j as (D x m2)
        const int64_t D = first[j], m1 = cur / D, m2 = numel[j] / D;
        CUDA_TRY(cudaMemcpyAsync(tj.p, host_t[j], (size_t)numel[j] * 16, cudaMemcpyHostToDevice, st));
        if ((rc = contract_svd_device(a, m1, tj.p, m2, D, er, b))) return rc;
        std::swap(a, b);
        cur = m1 * m2;
    }
    CUDA_TRY(cudaMemcpyAsync(host_out, a, (size_t)numel_out * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return QTN_OK;
}

// switch!(mps, i) (src/switch.jl:18-56) in one call: T = contract_svd(T1, T2, (ndims(T1), 1)) with er = 0, the two
// physical legs exchanged by permutedims, svd, T1' = U, T2' = diag(S) V'.  T1: (l1, 2, b) or (2, b) when l1 == 0;
// T2: (b, 2, r2) or (b, 2) when r2 == 0 (the boundary tensors of an MPS with two legs).
// host_u receives (l1, 2, bond) [(2, bond)], host_v (bond, 2, r2) [(bond, 2)]; *bond_out = min(2 l1, 2 r2).
int qtn_mps_switch_adjacent(const void* host_t1, int64_t l1, int64_t b, const void* host_t2, int64_t r2, void* host_u,
                            void* host_v, int64_t* bond_out) {
    QTN_API_GUARD();
    if (!host_t1 || !host_t2 || !host_u || !host_v || !bond_out) return fail(QTN_EINVAL, "qtn_mps_switch_adjacent: null argument");
    if (l1 < 0 || r2 < 0 || b < 1) return fail(QTN_EINVAL, "qtn_mps_switch_adjacent: bad extents");
    int rc = device_ready();
    if (rc) return rc;
    cudaStream_t st = stream();
    const int64_t L = std::max<int64_t>(l1, 1), R = std::max<int64_t>(r2, 1);
    const int64_t m1 = L * 2, m2 = 2 * R;
    DevBuf t1, t2, T, P, u, s, v, sv;
    const int64_t r = std::min(m1, m2);
    if ((rc = t1.alloc(m1 * b * 16)) || (rc = t2.alloc(b * m2 * 16)) || (rc = T.alloc(m1 * m2 * 16)) || (rc = P.alloc(m1 * m2 * 16)) ||
        (rc = u.alloc(m1 * r * 16)) || (rc = s.alloc(r * 8)) || (rc = v.alloc(r * m2 * 16)) || (rc = sv.alloc(r * m2 * 16)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(t1.p, host_t1, (size_t)m1 * b * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t2.p, host_t2, (size_t)b * m2 * 16, cudaMemcpyHostToDevice, st));
    if ((rc = contract_svd_device(t1.p, m1, t2.p, m2, b, 0.0, T.p))) return rc;   // src/switch.jl:26
    // T is (L, 2, 2, R): exchange the physical legs (src/switch.jl:28-36: [2,1,3] / [1,3,2] / [1,3,2,4] are this
    // permutation with the absent boundary legs dropped)
    const int64_t dims4[4] = {L, 2, 2, R};
    const int32_t perm4[4] = {1, 3, 2, 4};
    if ((rc = permutedims_device(T.p, 4, dims4, perm4, P.p))) return rc;
    SvdJob job{(double2*)P.p, m1, m2, (double2*)u.p, (double*)s.p, (double2*)v.p};
    int64_t k = 0;
    if ((rc = svd_batched_device(1, &job, -1.0, 0, &k, nullptr, nullptr))) return rc;   // src/switch.jl:39
    if ((rc = scale_copy((double2*)v.p, r, (double2*)sv.p, r, r, m2, (double*)s.p, 0))) return rc;   // src/switch.jl:41
    CUDA_TRY(cudaMemcpyAsync(host_u, u.p, (size_t)m1 * r * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(host_v, sv.p, (size_t)r * m2 * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *bond_out = r;
    return QTN_OK;
}

// ---------------- MPO x MPS (EXTENSION iii / iv) ------------------------------------------------------
static int check_mpo(const qtn_mps* m, const void* const* sites, const int64_t* dl, const int64_t* dr) {
    if (!m || !sites || !dl || !dr) return fail(QTN_EINVAL, "qtn_mps mpo: null argument");
    if (dl[0] != 1 || dr[m->n - 1] != 1) return fail(QTN_EINVAL, "MPO boundary bonds must be 1");
    for (int i = 0; i < m->n; ++i) {
        if (dl[i] < 1 || dr[i] < 1 || !sites[i]) return fail(QTN_EINVAL, "MPO site %d malformed", i + 1);
        if (i > 0 && dl[i] != dr[i - 1]) return fail(QTN_EINVAL, "MPO bonds of sites %d and %d do not match", i, i + 1);
    }
    return QTN_OK;
}

int qtn_mps_apply_mpo(qtn_mps* m, const void* const* host_mpo_sites, const int64_t* dl, const int64_t* dr, double er,
                      int64_t maxdim, double* disc_out) {
    QTN_API_GUARD();
    int rc = check_mpo(m, host_mpo_sites, dl, dr);
    if (rc) return rc;
    if ((rc = device_ready())) return rc;
    if (maxdim <= 0 || maxdim > m->cap) maxdim = m->cap;
    cudaStream_t st = stream();
    const int n = m->n;
    int64_t Dmax = 1, wtot = 0;
    for (int i = 0; i < n; ++i) { Dmax = std::max(Dmax, std::max(dl[i], dr[i])); wtot += dl[i] * 4 * dr[i]; }
    // 1. site-wise apply into fat sites (bonds multiply)
    std::vector<DevBuf> fat(n);
    std::vector<int64_t> lb(n), rb(n);
    DevBuf W;
    if ((rc = W.alloc(wtot * 16))) return rc;
    int64_t woff = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t L = m->lb[i], R = m->rb[i];
        lb[i] = L * dl[i];
        rb[i] = R * dr[i];
        if ((rc = fat[i].alloc((size_t)lb[i] * 2 * rb[i] * 16))) return rc;
        double2* w = (double2*)W.p + woff;
        CUDA_TRY(cudaMemcpyAsync(w, host_mpo_sites[i], (size_t)dl[i] * 4 * dr[i] * 16, cudaMemcpyHostToDevice, st));
        woff += dl[i] * 4 * dr[i];
        const int64_t tot = lb[i] * 2 * rb[i];
        mpo_apply_site_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, 148 * 8), 256, 0, st>>>(m->site[i], L, R, w, dl[i], dr[i], (double2*)fat[i].p);
        count_launch(1);
    }
    CUDA_TRY(cudaGetLastError());
    // 2. left-to-right sweep: orthogonalise (SVD without truncation), carry S*Vh to the right
    int64_t big = 0;
    for (int i = 0; i < n; ++i) big = std::max(big, lb[i] * 2 * rb[i]);
    DevBuf U, S, Vh, C, T;
    if ((rc = U.alloc(big * 16)) || (rc = S.alloc(std::max<int64_t>(big, 16) * 8)) || (rc = Vh.alloc(big * 16)) || (rc = C.alloc(big * 16)) ||
        (rc = T.alloc(big * 16)))
        return rc;
    for (int i = 0; i + 1 < n; ++i) {
        const int64_t mm = lb[i] * 2, nn = rb[i], r = std::min(mm, nn);
        const bool u_only = mm >= nn;  // carry = diag(S) V' = U^H M exactly: V is never accumulated
        if (u_only) {
            if ((rc = orth_columns(fat[i].p, T.p, C.p, mm, nn, U.p, (double*)S.p, nullptr))) return rc;
        } else {
            SvdJob job{(double2*)fat[i].p, mm, nn, (double2*)U.p, (double*)S.p, (double2*)Vh.p};
            int64_t k = 0;
            if ((rc = svd_batched_device(1, &job, -1.0, 0, &k, nullptr, nullptr))) return rc;
            if ((rc = scale_copy((double2*)U.p, mm, (double2*)fat[i].p, mm, mm, r, nullptr, 1))) return rc;  // site i <- U
        }
        // C = carry (r x nn): Q^H M, or diag(S) Vh after the full SVD; site i+1 <- C * site_{i+1} (nn x 2 rb[i+1])
        if (!u_only && (rc = scale_copy((double2*)Vh.p, r, (double2*)C.p, r, r, nn, (double*)S.p, 0))) return rc;
        const int64_t ncols = 2 * rb[i + 1];
        if ((rc = qtn_zgemm_device('N', 'N', r, ncols, nn, C.p, r, fat[i + 1].p, nn, T.p, r))) return rc;
        CUDA_TRY(cudaMemcpyAsync(fat[i + 1].p, T.p, (size_t)r * ncols * 16, cudaMemcpyDeviceToDevice, st));
        rb[i] = r;
        lb[i + 1] = r;
    }
    // 3. right-to-left sweep: truncate (er, maxdim); site i <- Vh[:k], site i-1 <- site_{i-1} * (U[:, :k] S).
    // The compressed sites go to temporaries and are committed to the handle only after the whole sweep succeeded:
    // a failure mid-sweep (memory, SVD non-convergence, capacity) leaves the MPS exactly as it was.
    std::vector<DevBuf> comp(n);
    std::vector<int64_t> new_lb(n), new_rb(n);
    for (int i = n - 1; i >= 1; --i) {
        const int64_t mm = lb[i], nn = 2 * rb[i], r = std::min(mm, nn);
        SvdJob job{(double2*)fat[i].p, mm, nn, (double2*)U.p, (double*)S.p, (double2*)Vh.p};
        int64_t k = 0;
        double disc = 0;
        if ((rc = svd_batched_device(1, &job, er, maxdim, &k, &disc, nullptr))) return rc;
        k = std::max<int64_t>(k, 1);
        if (disc_out) disc_out[i - 1] = disc;
        if ((size_t)k * nn > (size_t)2 * m->cap * m->cap) return fail(QTN_ENOMEM, "compressed site %d exceeds the MPS capacity", i + 1);
        if ((rc = comp[i].alloc((size_t)k * nn * 16))) return rc;
        if ((rc = scale_copy((double2*)Vh.p, r, (double2*)comp[i].p, k, k, nn, nullptr, 0))) return rc;     // (k, 2, rb)
        if ((rc = scale_copy((double2*)U.p, mm, (double2*)C.p, mm, mm, k, (double*)S.p, 1))) return rc;  // U[:, :k] S
        const int64_t rows = lb[i - 1] * 2;
        if ((rc = qtn_zgemm_device('N', 'N', rows, k, mm, fat[i - 1].p, rows, C.p, mm, T.p, rows))) return rc;
        CUDA_TRY(cudaMemcpyAsync(fat[i - 1].p, T.p, (size_t)rows * k * 16, cudaMemcpyDeviceToDevice, st));
        new_lb[i] = k;
        new_rb[i] = rb[i];
        rb[i - 1] = k;
    }
    if ((size_t)lb[0] * 2 * rb[0] > (size_t)2 * m->cap * m->cap) return fail(QTN_ENOMEM, "compressed site 1 exceeds the MPS capacity");
    CUDA_TRY(cudaStreamSynchronize(st));  // everything computed: commit
    for (int i = 0; i < n; ++i) {
        const void* src = i == 0 ? fat[0].p : comp[i].p;
        const int64_t l = i == 0 ? lb[0] : new_lb[i], r = i == 0 ? rb[0] : new_rb[i];
        CUDA_TRY(cudaMemcpyAsync(m->site[i], src, (size_t)l * 2 * r * 16, cudaMemcpyDeviceToDevice, st));
        m->lb[i] = l;
        m->rb[i] = r;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return QTN_OK;
}

// EXTENSION: orthonormal basis of the columns of a host matrix (m >= n); the gauge step of qtn_mps_apply_mpo.
int qtn_orth_columns(const void* host_a, int64_t m, int64_t n, void* host_q, int32_t* method_out) {
    QTN_API_GUARD();
    if (!host_a || !host_q) return fail(QTN_EINVAL, "qtn_orth_columns: null argument");
    if (n < 1 || m < n) return fail(QTN_EINVAL, "qtn_orth_columns: needs m >= n >= 1 (got %lld x %lld)", (long long)m, (long long)n);
    int rc = device_ready();
    if (rc) return rc;
    DevBuf M, keep, U, S, carry;
    const size_t bytes = (size_t)m * n * 16;
    if ((rc = M.alloc(bytes)) || (rc = keep.alloc(bytes)) || (rc = U.alloc(bytes)) || (rc = S.alloc((size_t)n * 8)) ||
        (rc = carry.alloc((size_t)n * n * 16)))
        return rc;
    cudaStream_t st = stream();
    CUDA_TRY(cudaMemcpyAsync(M.p, host_a, bytes, cudaMemcpyHostToDevice, st));
    int method = 0;
    if ((rc = orth_columns(M.p, keep.p, carry.p, m, n, U.p, (double*)S.p, &method))) return rc;
    CUDA_TRY(cudaMemcpyAsync(host_q, M.p, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (method_out) *method_out = method;
    return QTN_OK;
}

int qtn_mps_expect_mpo(const qtn_mps* m, const void* const* host_mpo_sites, const int64_t* dl, const int64_t* dr, double out[2]) {
    QTN_API_GUARD();
    int rc = check_mpo(m, host_mpo_sites, dl, dr);
    if (rc) return rc;
    if (!out) return fail(QTN_EINVAL, "null argument");
    if ((rc = device_ready())) return rc;
    if (m->lb[0] != 1 || m->rb[m->n - 1] != 1) return fail(QTN_EINVAL, "qtn_mps_expect_mpo: boundary bonds must be 1");
    cudaStream_t st = stream();
    int64_t Dmax = 1;
    for (int i = 0; i < m->n; ++i) Dmax = std::max(Dmax, std::max(dl[i], dr[i]));
    const int64_t cap = m->cap;
    DevBuf E, E2, T1, T2, W;
    if ((rc = E.alloc(cap * Dmax * cap * 16)) || (rc = E2.alloc(cap * Dmax * cap * 16)) || (rc = T1.alloc(cap * Dmax * 2 * cap * 16)) ||
        (rc = T2.alloc(cap * 2 * Dmax * cap * 16)) || (rc = W.alloc(Dmax * 4 * Dmax * 16)))
        return rc;
    const double2 one = make_double2(1.0, 0.0);
    CUDA_TRY(cudaMemcpyAsync(E.p, &one, 16, cudaMemcpyHostToDevice, st));
    void *e = E.p, *e2 = E2.p;
    for (int i = 0; i < m->n; ++i) {
        const int64_t L = m->lb[i], R = m->rb[i];
        CUDA_TRY(cudaMemcpyAsync(W.p, host_mpo_sites[i], (size_t)dl[i] * 4 * dr[i] * 16, cudaMemcpyHostToDevice, st));
        // T1[(la,a), (p,rb)] = E[(la,a), lb] A[lb, (p,rb)]
        if ((rc = qtn_zgemm_device('N', 'N', L * dl[i], 2 * R, L, e, L * dl[i], m->site[i], L, T1.p, L * dl[i]))) return rc;
        const int64_t tot = L * 2 * dr[i] * R;
        env_apply_w_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, 148 * 8), 256, 0, st>>>((const double2*)T1.p, L, R, (const double2*)W.p, dl[i], dr[i], (double2*)T2.p);
        count_launch(1);
        // E'[ra, (b,rb)] = A^H[ra, (la,p')] T2[(la,p'), (b,rb)]
        if ((rc = qtn_zgemm_device('C', 'N', R, dr[i] * R, 2 * L, m->site[i], 2 * L, T2.p, 2 * L, e2, R))) return rc;
        std::swap(e, e2);
    }
    double2 res;
    CUDA_TRY(cudaMemcpyAsync(&res, e, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    out[0] = res.x;
    out[1] = res.y;
    return QTN_OK;
}

}  // extern "C"
