"""Emulation: does restricting the eigensolve of a block-pair visit to CROSS rotations (p in block I, q in block J:
16 tournament steps instead of 31) change the outer sweep count of the blocked one-sided Jacobi?  Within-block pairs
are then rotated only in round 0 of every sweep (full 31-step visit).  theta-like matrices of the MPS path."""
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


def rotate_pairs(G, W, pairs, tol):
    """Apply the disjoint rotations of `pairs` (list of (p, q)) to G two-sidedly and to W; returns #rotations."""
    n = G.shape[0]
    J = np.eye(n, dtype=complex)
    nrot = 0
    for p, q in pairs:
        g = G[p, q]
        ag = abs(g)
        a, b = G[p, p].real, G[q, q].real
        if ag * ag <= tol * tol * abs(a) * abs(b) or ag == 0.0:
            if a < b:
                J[:, [p, q]] = J[:, [q, p]] * np.array([1, -1])   # pure swap (sort)
            continue
        nrot += 1
        ph = g / ag
        zeta = (b - a) / (2 * ag)
        t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
        c = 1 / np.sqrt(1 + t * t)
        s = c * t
        if a - t * ag < b + t * ag:
            c, s = s, -c
        J[p, p] = c; J[p, q] = s * ph; J[q, p] = -s * np.conj(ph); J[q, q] = c
    G[:] = J.conj().T @ G @ J
    W[:] = W @ J
    return nrot


def visit(G, tol, cross_only, b):
    n = 2 * b
    W = np.eye(n, dtype=complex)
    nrot = 0
    if cross_only:
        for step in range(b):
            nrot += rotate_pairs(G, W, [(k, b + (k + step) % b) for k in range(b)], tol)
    else:
        for step in range(n - 1):
            pairs = []
            for k in range(b):
                if k == 0:
                    p, q = n - 1, step
                else:
                    p, q = (step + k) % (n - 1), (step - k + (n - 1)) % (n - 1)
                pairs.append((min(p, q), max(p, q)))
            nrot += rotate_pairs(G, W, pairs, tol)
    return W, nrot


def block_jacobi(A, b, mode, tol=1.6e-14, max_sweeps=40):
    A = A.copy()
    n = A.shape[1]
    nb = n // b
    sweeps = 0
    hist = []
    while sweeps < max_sweeps:
        total = 0
        for rnd in range(nb - 1):
            for k in range(nb // 2):
                if k == 0:
                    i, j = nb - 1, rnd
                else:
                    i, j = (rnd + k) % (nb - 1), (rnd - k + (nb - 1)) % (nb - 1)
                i, j = min(i, j), max(i, j)
                cols = np.r_[i * b:(i + 1) * b, j * b:(j + 1) * b]
                P = A[:, cols]
                G = P.conj().T @ P
                cross = mode == "cross" and rnd >= 1
                W, nrot = visit(G, tol, cross, b)
                total += nrot
                A[:, cols] = P @ W
        sweeps += 1
        hist.append(total)
        if total == 0:
            break
    return sweeps, hist, np.sort(np.linalg.norm(A, axis=0))[::-1]


if __name__ == "__main__":
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rng = np.random.default_rng(1)
    th = theta_like(chi, rng)
    sref = np.linalg.svd(th, compute_uv=False)
    print("theta-like %dx%d" % th.shape)
    for mode in ("full", "cross"):
        t0 = time.time()
        sw, hist, s = block_jacobi(th, 16, mode)
        print("mode=%-5s outer sweeps %2d  rotations/sweep %s  max sigma err %.1e (%.0f s)" % (mode, sw, hist, np.abs(s - sref).max() / sref[0], time.time() - t0), flush=True)
