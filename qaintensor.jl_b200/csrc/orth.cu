// EXTENSION (no reference counterpart): orthonormalisation of a tall site matrix for the gauge sweep of
// MPO x MPS compression (cfg 5).  The reference would use `svd` here too (cf. src/mps.jl:63); a gauge sweep
// needs no singular values, so the columns are orthonormalised by blocked CholeskyQR2, which is GEMM work
// (FP64 DMMA kernel) apart from one 64 x 64 tile factorisation per block column:
//
//   twice:  G = X^H X  ->  G = R^H R (blocked right-looking Cholesky)  ->  X <- X R^-1
//
// Two passes give |Q^H Q - I| = O(eps) while cond(X) < ~1e8.  The result is verified on the device, NaN-aware:
// max |Q^H Q - I| (Q is orthonormal) and max |Q (Q^H M) - M| (Q spans the columns of M; a numerically
// rank-deficient M can pass the first test with a wrong span).  When either check fails the caller falls
// back to the Jacobi SVD, so an ill-conditioned or rank-deficient input costs time, never accuracy.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "kernels.cuh"
#include "qtn_internal.h"

namespace qtn {
cudaStream_t stream();
void count_launch(int64_t n);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(QTN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

namespace {
constexpr int kNb = 64;          // block column width
constexpr int kPitch = kNb + 1;  // shared-memory row pitch (elements)

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// One CTA.  In place upper Cholesky of the Hermitian b x b tile at G (only its upper triangle is read):
// G <- R with a zeroed strict lower triangle; X <- R^-1 and NX <- -R^-1 as dense 64 x 64 column-major tiles
// (zero outside b x b).  A non-positive pivot yields NaNs, which the final orthogonality check catches.
__global__ void __launch_bounds__(256) potrf_tile_kernel(double2* __restrict__ G, int64_t ld, int b, double2* __restrict__ X,
                                                         double2* __restrict__ NX) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* T = reinterpret_cast<double2*>(smem_raw);  // T[r * kPitch + c]
    double2* Xi = T + kNb * kPitch;
    const int tid = threadIdx.x;
    for (int e = tid; e < kNb * kNb; e += blockDim.x) {
        const int r = e % kNb, c = e / kNb;
        T[r * kPitch + c] = (r < b && c < b) ? G[(int64_t)c * ld + r] : make_double2(0.0, 0.0);
        Xi[r * kPitch + c] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    for (int k = 0; k < b; ++k) {
        const double d = sqrt(T[k * kPitch + k].x);
        const double inv = 1.0 / d;
        __syncthreads();
        if (tid < b - k) {
            const int j = k + tid;
            double2 v = T[k * kPitch + j];
            T[k * kPitch + j] = (tid == 0) ? make_double2(d, 0.0) : make_double2(v.x * inv, v.y * inv);
        }
        __syncthreads();
        const int nr = b - k - 1;
        for (int e = tid; e < nr * nr; e += blockDim.x) {
            const int i = k + 1 + e / nr, j = k + 1 + e % nr;
            if (j >= i) {
                const double2 a = T[k * kPitch + i], c = T[k * kPitch + j];  // T[i][j] -= conj(R[k][i]) R[k][j]
                double2 t = T[i * kPitch + j];
                t.x -= a.x * c.x + a.y * c.y;
                t.y -= a.x * c.y - a.y * c.x;
                T[i * kPitch + j] = t;
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < b * b; e += blockDim.x) {
        const int r = e % b, c = e / b;
        if (r > c) T[r * kPitch + c] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    // row i of R^-1 depends only on its own earlier entries: one thread per row, no barriers
    if (tid < b) {
        const int i = tid;
        for (int j = i; j < b; ++j) {
            const double rjj = 1.0 / T[j * kPitch + j].x;
            if (j == i) {
                Xi[i * kPitch + j] = make_double2(rjj, 0.0);
            } else {
                double2 acc = make_double2(0.0, 0.0);
                for (int k = i; k < j; ++k) {
                    const double2 p = cmul(Xi[i * kPitch + k], T[k * kPitch + j]);
                    acc.x += p.x;
                    acc.y += p.y;
                }
                Xi[i * kPitch + j] = make_double2(-acc.x * rjj, -acc.y * rjj);
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < kNb * kNb; e += blockDim.x) {
        const int r = e % kNb, c = e / kNb;
        const double2 x = Xi[r * kPitch + c];
        X[e] = x;
        NX[e] = make_double2(-x.x, -x.y);
        if (r < b && c < b) G[(int64_t)c * ld + r] = T[r * kPitch + c];
    }
}

// dst[r, c] = src[r, c], neg[r, c] = -src[r, c] for a rows x cols block (separate leading dimensions)
__global__ void copy_neg_kernel(const double2* __restrict__ src, int64_t lds, int64_t rows, int64_t cols, double2* __restrict__ dst,
                                int64_t ldd, double2* __restrict__ neg, int64_t ldn) {
    const int64_t tot = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e % rows, c = e / rows;
        const double2 v = src[c * lds + r];
        dst[c * ldd + r] = v;
        neg[c * ldn + r] = make_double2(-v.x, -v.y);
    }
}

// *out = max over the n x n matrix of |G - I| (max of the real and imaginary magnitudes); NaN counts as +inf.
// Non-negative doubles order like their bit patterns, so the reduction is an integer atomicMax.
__global__ void orth_defect_kernel(const double2* __restrict__ G, int64_t n, unsigned long long* __restrict__ out) {
    double worst = 0.0;
    const int64_t tot = n * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e % n, c = e / n;
        const double2 v = G[e];
        double d = fmax(fabs(v.x - (r == c ? 1.0 : 0.0)), fabs(v.y));
        if (!(d == d)) d = INFINITY;  // fmax drops NaNs: test both parts explicitly
        if (!(v.x == v.x) || !(v.y == v.y)) d = INFINITY;
        worst = fmax(worst, d);
    }
    for (int o = 16; o; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(worst));
}

// out[0] = max |A - B| over n elements, out[1] = max |B| (real / imaginary magnitudes); NaN counts as +inf.
__global__ void residual_kernel(const double2* __restrict__ A, const double2* __restrict__ B, int64_t n, unsigned long long* __restrict__ out) {
    double worst = 0.0, scale = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double2 a = A[e], b = B[e];
        double d = fmax(fabs(a.x - b.x), fabs(a.y - b.y));
        if (!(a.x == a.x) || !(a.y == a.y)) d = INFINITY;
        worst = fmax(worst, d);
        scale = fmax(scale, fmax(fabs(b.x), fabs(b.y)));
    }
    for (int o = 16; o; o >>= 1) {
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, (unsigned long long)__double_as_longlong(worst));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(scale));
    }
}

int grid_for(int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, 148 * 8)); }

using Scratch = PoolBuf;   // workspace pool (pool.cu)

// G (n x n, ld n) <- upper Cholesky factor in its upper block rows (strictly lower blocks are left stale);
// dinv / ndinv <- +-R_jj^-1 per diagonal block; work = 2 * 64 * n elements.
int chol_upper_blocked(double2* G, int64_t n, double2* dinv, double2* ndinv, double2* work) {
    cudaStream_t st = stream();
    const size_t smem = (size_t)2 * kNb * kPitch * sizeof(double2);
    static bool attr = false;
    if (!attr) { CUDA_TRY(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    double2* rtmp = work;             // 64 x rest, ld 64
    double2* nbuf = work + kNb * n;   // its negative
    int rc;
    for (int64_t j0 = 0, jb = 0; j0 < n; j0 += kNb, ++jb) {
        const int b = (int)std::min<int64_t>(kNb, n - j0);
        double2* diag = G + j0 * n + j0;
        potrf_tile_kernel<<<1, 256, smem, st>>>(diag, n, b, dinv + jb * kNb * kNb, ndinv + jb * kNb * kNb);
        CUDA_TRY(cudaGetLastError());
        count_launch(1);
        const int64_t rest = n - j0 - b;
        if (rest <= 0) break;
        double2* row = G + (j0 + b) * n + j0;  // G[j0 : j0+b, j0+b :]
        // R_j,rest = R_jj^-H G_j,rest
        if ((rc = zgemm_dense('C', 'N', b, rest, b, dinv + jb * kNb * kNb, kNb, row, n, rtmp, kNb, false))) return rc;
        copy_neg_kernel<<<grid_for((int64_t)b * rest), 256, 0, st>>>(rtmp, kNb, b, rest, row, n, nbuf, kNb);
        CUDA_TRY(cudaGetLastError());
        count_launch(1);
        // G_rest,rest -= R_j,rest^H R_j,rest
        if ((rc = zgemm_dense('C', 'N', rest, rest, b, nbuf, kNb, rtmp, kNb, G + (j0 + b) * n + (j0 + b), n, true))) return rc;
    }
    return QTN_OK;
}

// X (n x n, zero-initialised) <- R^-1 for the upper-triangular R held in the upper block rows of G.
int rinv_upper_blocked(const double2* G, int64_t n, const double2* dinv, const double2* ndinv, double2* X, double2* work) {
    cudaStream_t st = stream();
    int rc;
    for (int64_t j0 = 0, jb = 0; j0 < n; j0 += kNb, ++jb) {
        const int b = (int)std::min<int64_t>(kNb, n - j0);
        CUDA_TRY(cudaMemcpy2DAsync(X + j0 * n + j0, (size_t)n * 16, dinv + jb * kNb * kNb, (size_t)kNb * 16, (size_t)b * 16, (size_t)b,
                                   cudaMemcpyDeviceToDevice, st));
        if (j0 == 0) continue;
        // X[:j0, jblk] = -(X[:j0, :j0] R[:j0, jblk]) R_jj^-1
        if ((rc = zgemm_dense('N', 'N', j0, b, j0, X, n, G + j0 * n, n, work, j0, false))) return rc;
        if ((rc = zgemm_dense('N', 'N', j0, b, b, work, j0, ndinv + jb * kNb * kNb, kNb, X + j0 * n, n, false))) return rc;
    }
    return QTN_OK;
}
}  // namespace

int orth_cholqr2(void* dev_m, const void* dev_keep, void* dev_c, int64_t m, int64_t n, bool* ok) {
    *ok = false;
    if (m < n || n < 1) return fail(QTN_EINVAL, "orth_cholqr2: needs m >= n >= 1 (got %lld x %lld)", (long long)m, (long long)n);
    cudaStream_t st = stream();
    const int64_t nblk = (n + kNb - 1) / kNb;
    Scratch G, X, D, W, Q2, defect;
    int rc;
    if ((rc = G.alloc((size_t)n * n * 16)) || (rc = X.alloc((size_t)n * n * 16)) || (rc = D.alloc((size_t)2 * nblk * kNb * kNb * 16)) ||
        (rc = W.alloc((size_t)2 * kNb * n * 16)) || (rc = Q2.alloc((size_t)m * n * 16)) || (rc = defect.alloc(32)))
        return rc;
    double2* dinv = (double2*)D.p;
    double2* ndinv = dinv + nblk * kNb * kNb;
    double2* src = (double2*)dev_m;
    double2* dst = (double2*)Q2.p;
    for (int pass = 0; pass < 2; ++pass) {
        if ((rc = zgemm_dense('C', 'N', n, n, m, src, m, src, m, G.p, n, false))) return rc;
        if ((rc = chol_upper_blocked((double2*)G.p, n, dinv, ndinv, (double2*)W.p))) return rc;
        CUDA_TRY(cudaMemsetAsync(X.p, 0, (size_t)n * n * 16, st));
        if ((rc = rinv_upper_blocked((const double2*)G.p, n, dinv, ndinv, (double2*)X.p, (double2*)W.p))) return rc;
        if ((rc = zgemm_dense('N', 'N', m, n, n, src, m, X.p, n, dst, m, false))) return rc;
        std::swap(src, dst);
    }
    // two swaps: Q is back in dev_m.  Verify orthonormality and the span.
    unsigned long long* dd = (unsigned long long*)defect.p;
    CUDA_TRY(cudaMemsetAsync(dd, 0, 32, st));
    if ((rc = zgemm_dense('C', 'N', n, n, m, dev_m, m, dev_m, m, G.p, n, false))) return rc;
    orth_defect_kernel<<<grid_for(n * n), 256, 0, st>>>((const double2*)G.p, n, dd);
    CUDA_TRY(cudaGetLastError());
    if ((rc = zgemm_dense('C', 'N', n, n, m, dev_m, m, dev_keep, m, dev_c, n, false))) return rc;   // C = Q^H M
    if ((rc = zgemm_dense('N', 'N', m, n, n, dev_m, m, dev_c, n, Q2.p, m, false))) return rc;      // Q C
    residual_kernel<<<grid_for(m * n), 256, 0, st>>>((const double2*)Q2.p, (const double2*)dev_keep, m * n, dd + 1);
    CUDA_TRY(cudaGetLastError());
    count_launch(2);
    double h[3] = {0.0, 0.0, 0.0};  // orthogonality defect, residual, scale of M
    CUDA_TRY(cudaMemcpyAsync(h, dd, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const double tol = 2e-14 * std::sqrt((double)m);  // ~100x what a well-conditioned input leaves
    *ok = h[0] <= tol && h[1] <= tol * h[2];
    return QTN_OK;
}

}  // namespace qtn
