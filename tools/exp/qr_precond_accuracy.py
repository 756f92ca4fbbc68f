"""Accuracy of the Cholesky-preconditioned U-only SVD data flow proposed in DESIGN.md section 7 (numpy emulation).

  G = theta^H theta,  R = chol(G)  (optionally refined: R <- chol((theta R^-1)^H (theta R^-1)) R, i.e. CholeskyQR2),
  X = R^H,  one-sided Jacobi from the right:  X W = Y  (columns orthogonal, W never accumulated),
  sigma_j = ||y_j||,   S V^H = Y^H  (free),   U = theta (Y Sigma^-1) Sigma^-1  (one GEMM + scaling, kept columns only).

Checks against LAPACK on theta-like and on graded matrices: kept singular values (bar 1e-10 sigma_max), orthogonality of U_k,
|U_k (S V^H)_k - theta_k(LAPACK)| and the fidelity of the truncated state (bar 1e-9).  numpy's own SVD of X stands in for
the Jacobi sweeps (same backward-stable result class); the question here is what the Cholesky / recovery steps cost."""
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from block_jacobi_sweeps import theta_like  # noqa: E402


def precond_svd(theta, k, passes):
    G = theta.conj().T @ theta
    R = sla.cholesky(G, lower=False)
    if passes == 2:
        Q1 = sla.solve_triangular(R, theta.conj().T, trans="C", lower=False).conj().T  # theta R^-1
        R = sla.cholesky(Q1.conj().T @ Q1, lower=False) @ R
    X = R.conj().T
    Ux, s, Wh = np.linalg.svd(X)          # stands in for one-sided Jacobi: Y = X W = Ux diag(s)
    Y = Ux * s
    svh = Y.conj().T[:k]                  # S V^H, rows = kept
    U = (theta @ Ux[:, :k]) / s[:k]
    return U, s, svh


def report(name, theta, k):
    U0, s0, Vh0 = np.linalg.svd(theta)
    ref = (U0[:, :k] * s0[:k]) @ Vh0[:k]
    for passes in (1, 2):
        try:
            U, s, svh = precond_svd(theta, k, passes)
        except np.linalg.LinAlgError as e:
            print("%-34s chol passes %d: Cholesky failed (%s)" % (name, passes, e))
            continue
        approx = U @ svh
        fid = abs(np.vdot(ref, approx)) / (np.linalg.norm(ref) * np.linalg.norm(approx))
        print("%-34s chol passes %d: sigma err %.1e (of sigma_max), |U^H U - I| %.1e, |theta_k - ref| %.1e, 1 - fidelity %.1e"
              % (name, passes, np.abs(s[:k] - s0[:k]).max() / s0[0], np.abs(U.conj().T @ U - np.eye(k)).max(),
                 np.abs(approx - ref).max() / s0[0], 1 - fid))


if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rng = np.random.default_rng(2)
    th = theta_like(chi, rng)
    report("theta-like %d (cond 2e3)" % (2 * chi), th, chi)
    n = 2 * chi
    for decay in (1e-4, 1e-6, 1e-8):
        u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        s = decay ** (np.arange(n) / (n - 1))
        k = int(np.searchsorted(-s, -1e-5))  # keep sigma >= 1e-5 sigma_max, as a 1e-10 tail-weight cutoff does
        report("graded, sigma_min/max %.0e, k=%d" % (decay, k), (u * s) @ v.conj().T, max(k, 1))
