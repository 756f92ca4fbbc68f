"""Oracle restatement of ``src/svd.jl`` and the SVD idioms around it (test infrastructure).

``LinearAlgebra.svd`` (Julia stdlib -> LAPACK ``zgesdd``, thin) is restated with
``scipy.linalg.svd(full_matrices=False, lapack_driver="gesdd")`` -- the same
LAPACK routine.  Julia's ``U, S, V = svd(A)`` returns ``V`` (not ``V'``); scipy
returns ``Vh = V'``.
"""
import numpy as np
import scipy.linalg


def svd(A):
    """Thin SVD, returns (U, S, Vh) with Vh = adjoint(V) of the reference's V."""
    A = np.asarray(A)
    if A.size == 0:
        k = min(A.shape)
        return (np.zeros((A.shape[0], k), A.dtype), np.zeros(k), np.zeros((k, A.shape[1]), A.dtype))
    return scipy.linalg.svd(A, full_matrices=False, lapack_driver="gesdd")


def truncation_rank(S, er=0.0, maxdim=None):
    """Number of singular values kept by the rule of ``src/svd.jl:29-33``.

    ``tail[r] = sqrt(S[n]^2 + ... + S[n-r+1]^2)`` (cumsum of the reversed
    squares); ``r* = first r with tail[r] > er`` (strict); ``k = n - r* + 1``.
    If no tail exceeds ``er`` the reference's ``findfirst`` returns ``nothing``
    and it throws; the documented extension here is k = 0.
    EXTENSION (no reference counterpart): ``k <- min(k, maxdim)`` after the cutoff.
    """
    S = np.asarray(S, dtype=np.float64)
    n = len(S)
    tail = np.sqrt(np.cumsum(S[::-1] ** 2))
    hit = np.nonzero(tail > er)[0]
    k = 0 if len(hit) == 0 else n - (int(hit[0]) + 1) + 1
    if maxdim is not None:
        k = min(k, int(maxdim))
    return k


def contract_svd(T1, T2, indx, er=0.0):
    """``contract_svd(T1, T2, (i1, i2); er)`` (src/svd.jl:7-38); arrays in, array out."""
    if not er >= 0:
        raise ValueError("Error must be positive")
    T1 = np.asarray(T1)
    T2 = np.asarray(T2)
    i1, i2 = indx
    n1, n2 = T1.ndim, T2.ndim
    # Julia's size(A, d) is 1 for d > ndims(A) (test/test_svd.jl:47-48 relies on it)
    D1 = T1.shape[i1 - 1] if i1 <= n1 else 1
    D2 = T2.shape[i2 - 1] if i2 <= n2 else 1
    if D1 != D2:
        raise ValueError("Dimensions of contraction legs do not match")
    newdim = list(T1.shape[:i1 - 1]) + list(T1.shape[i1:]) + list(T2.shape[:i2 - 1]) + list(T2.shape[i2:])
    p1 = [a for a in range(n1) if a != i1 - 1] + [i1 - 1]
    p2 = [i2 - 1] + [a for a in range(n2) if a != i2 - 1]
    T1p = np.reshape(np.transpose(T1, p1), (-1, D1), order="F")
    T2p = np.reshape(np.transpose(T2, p2), (D1, -1), order="F")
    U1, S1, V1h = svd(T1p)
    U2, S2, V2h = svd(T2p)
    k1 = truncation_rank(S1, er)
    k2 = truncation_rank(S2, er)
    if k1 == 0 or k2 == 0:
        raise ValueError("truncation removed every singular value (reference: findfirst -> nothing)")
    T = (U1[:, :k1] * S1[:k1]) @ V1h[:k1, :] @ (U2[:, :k2] * S2[:k2]) @ V2h[:k2, :]
    return np.reshape(T, newdim, order="F")


def svd_split(theta, er=0.0, maxdim=None):
    """``U, S, V = svd(theta)``; left = U[:, :k], right = diagm(S)*V' rows :k.

    The split used by ``src/switch.jl:39-52`` / ``src/mps.jl:63-76`` (there with
    k = all); truncation by ``truncation_rank`` is the documented extension.
    Returns (left, right, S_kept, discarded_weight).
    """
    U, S, Vh = svd(theta)
    k = truncation_rank(S, er, maxdim) if (er > 0 or maxdim is not None) else len(S)
    k = max(k, 1)
    disc = float(np.sqrt(np.sum(S[k:] ** 2)))
    return U[:, :k], S[:k, None] * Vh[:k, :], S[:k], disc
