"""numpy emulation of the blocked CholeskyQR2 orthogonalisation planned for csrc/orth.cu (block algebra only)."""
import numpy as np


def chol_tile(T):
    n = T.shape[0]
    T = T.copy()
    R = np.zeros_like(T)
    for k in range(n):
        d = np.sqrt(T[k, k].real)
        R[k, k] = d
        R[k, k + 1:] = T[k, k + 1:] / d
        for i in range(k + 1, n):
            T[i, i:] -= np.conj(R[k, i]) * R[k, i:]
    X = np.zeros_like(T)
    for i in range(n):
        for j in range(i, n):
            if i == j:
                X[i, j] = 1.0 / R[j, j]
            else:
                X[i, j] = -(X[i, i:j] @ R[i:j, j]) / R[j, j]
    return R, X


def chol_blocked(G, nb):
    n = G.shape[0]
    G = G.copy()
    dinv = []
    for j0 in range(0, n, nb):
        b = min(nb, n - j0)
        Rjj, Xjj = chol_tile(G[j0:j0 + b, j0:j0 + b])
        G[j0:j0 + b, j0:j0 + b] = Rjj
        dinv.append(Xjj)
        if j0 + b < n:
            Rjr = Xjj.conj().T @ G[j0:j0 + b, j0 + b:]
            G[j0:j0 + b, j0 + b:] = Rjr
            G[j0 + b:, j0 + b:] += (-Rjr).conj().T @ Rjr
    return G, dinv          # upper block rows of G hold R; strictly-lower blocks are stale


def rinv_blocked(R, dinv, nb):
    n = R.shape[0]
    X = np.zeros_like(R)
    for jb, j0 in enumerate(range(0, n, nb)):
        b = min(nb, n - j0)
        X[j0:j0 + b, j0:j0 + b] = dinv[jb]
        if j0:
            tmp = X[:j0, :j0] @ R[:j0, j0:j0 + b]
            X[:j0, j0:j0 + b] = tmp @ (-dinv[jb])
    return X


def cholqr2(M, nb=64):
    Q = M
    for _ in range(2):
        G = Q.conj().T @ Q
        R, dinv = chol_blocked(G, nb)
        Q = Q @ rinv_blocked(R, dinv, nb)
    return Q


rng = np.random.default_rng(0)
m, n = 384, 192
for cond in (1e1, 1e3, 1e6, 1e7, 1e8, 1e10):
    U = np.linalg.qr(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))[0]
    V = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))[0]
    M = (U * np.logspace(0, -np.log10(cond), n)) @ V.conj().T
    with np.errstate(all="ignore"):
        Q = cholqr2(M, 64)
        orth = np.max(np.abs(Q.conj().T @ Q - np.eye(n)))
        C = Q.conj().T @ M
        rec = np.max(np.abs(Q @ C - M)) / np.max(np.abs(M))
    print(f"cond {cond:.0e}: |Q^H Q - I|max = {orth:.2e}   |Q Q^H M - M|/|M| = {rec:.2e}")
