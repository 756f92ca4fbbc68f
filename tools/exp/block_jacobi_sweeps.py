"""Outer-sweep counts of the BLOCKED one-sided Jacobi of csrc/svd.cu under different inner solvers (numpy emulation).

The kernel pairs column blocks (circle method), forms the Gram G = P^H P of a 2b-column panel, runs `inner` cyclic
two-sided Jacobi sweeps on G and applies the accumulated W to the panel.  This script counts the outer sweeps to
convergence (no pair above tol in a whole sweep) for
  inner = 1, 2, 3 cyclic eigen-sweeps per visit, and `full` (exact eigendecomposition of G per visit),
  block sizes b = 16 and 32,
on theta-like matrices of the MPS path (two random site tensors contracted with a Haar 2-qubit gate) so that the
trade-off "cheaper visits vs fewer sweeps" can be judged without GPU time.  Usage: python block_jacobi_sweeps.py [chi]"""
import sys
import time

import numpy as np


def theta_like(chi, rng):
    a = rng.standard_normal((chi, 2, chi)) + 1j * rng.standard_normal((chi, 2, chi))
    b = rng.standard_normal((chi, 2, chi)) + 1j * rng.standard_normal((chi, 2, chi))
    z = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    g, _ = np.linalg.qr(z)
    th = np.einsum("lsb,btr->lstr", a, b)
    th = np.einsum("uvst,lstr->luvr", g.reshape(2, 2, 2, 2), th)
    th = th.reshape(2 * chi, 2 * chi)
    return th / np.linalg.norm(th)


def inner_jacobi(G, sweeps, tol):
    """`sweeps` cyclic two-sided Jacobi sweeps on the Hermitian G; returns (W, rotated?)."""
    n = G.shape[0]
    G = G.copy()
    W = np.eye(n, dtype=complex)
    rotated = False
    for _ in range(sweeps):
        for p in range(n - 1):
            for q in range(p + 1, n):
                g = G[p, q]
                ag = abs(g)
                a, b = G[p, p].real, G[q, q].real
                if ag * ag <= tol * tol * abs(a) * abs(b) or ag == 0.0:
                    continue
                rotated = True
                ph = g / ag
                zeta = (b - a) / (2 * ag)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
                c = 1 / np.sqrt(1 + t * t)
                s = c * t
                J = np.array([[c, s * ph], [-s * np.conj(ph), c]])
                G[:, [p, q]] = G[:, [p, q]] @ J
                G[[p, q], :] = J.conj().T @ G[[p, q], :]
                W[:, [p, q]] = W[:, [p, q]] @ J
    return W, rotated


def block_jacobi(A, b, inner, tol=1.6e-14, max_sweeps=60):
    A = A.copy()
    n = A.shape[1]
    nb = n // b
    order = list(range(nb))
    sweeps = 0
    while sweeps < max_sweeps:
        any_rot = False
        ring = order[:]
        for _ in range(nb - 1):
            for k in range(nb // 2):
                i, j = ring[k], ring[nb - 1 - k]
                cols = np.r_[i * b:(i + 1) * b, j * b:(j + 1) * b]
                P = A[:, cols]
                G = P.conj().T @ P
                d = np.sqrt(np.abs(np.diag(G).real))
                off = np.abs(G) - np.diag(np.abs(np.diag(G)))
                if not (off > tol * np.outer(d, d)).any():
                    continue
                any_rot = True
                if inner == "full":
                    w, V = np.linalg.eigh(G)
                    W = V[:, ::-1]
                else:
                    W, _ = inner_jacobi(G, inner, tol)
                A[:, cols] = P @ W
            ring = [ring[0]] + [ring[-1]] + ring[1:-1]
        sweeps += 1
        if not any_rot:
            break
    return sweeps, np.sort(np.linalg.norm(A, axis=0))[::-1]


if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rng = np.random.default_rng(1)
    th = theta_like(chi, rng)
    sref = np.linalg.svd(th, compute_uv=False)
    print("theta-like %dx%d, sigma_max/sigma_min = %.2e" % (th.shape[0], th.shape[1], sref[0] / sref[-1]))
    for b in (16, 32):
        for inner in (1, 2, 3, "full"):
            t0 = time.time()
            sw, s = block_jacobi(th, b, inner)
            print("b=%2d inner=%-4s outer sweeps %2d  max sigma err %.1e  (%.0f s)" % (b, inner, sw, np.abs(s - sref).max() / sref[0], time.time() - t0),
                  flush=True)
