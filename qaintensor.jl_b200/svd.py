"""Truncated SVD and ``contract_svd`` (src/svd.jl:7-38) through the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import arr_i64, as_c128, check, lib


def svd_trunc(A, er=-1.0, maxdim=0):
    """Thin SVD on the GPU with the tail-norm rule of src/svd.jl:29-33 and the
    ``maxdim`` cap (EXTENSION).  Returns (U, S, Vh, k): full thin factors and the number
    of kept values.  ``er < 0`` keeps everything (plain ``LinearAlgebra.svd``)."""
    _lib.require_device()
    A = as_c128(A)
    if A.ndim != 2:
        raise ValueError("svd_trunc expects a matrix")
    m, n = A.shape
    r = min(m, n)
    U = np.zeros((m, r), dtype=np.complex128, order="F")
    S = np.zeros(max(r, 1), dtype=np.float64)
    Vh = np.zeros((r, n), dtype=np.complex128, order="F")
    k = C.c_int64(0)
    check(lib.qtn_svd_trunc(A.ctypes.data_as(C.c_void_p), m, n, float(er), int(maxdim), U.ctypes.data_as(C.c_void_p),
                            S.ctypes.data_as(C.POINTER(C.c_double)), Vh.ctypes.data_as(C.c_void_p), C.byref(k)))
    return U, S[:r], Vh, int(k.value)


def svd(A):
    """``U, S, V = svd(A)`` of the reference, returned as (U, S, Vh = V')."""
    U, S, Vh, _ = svd_trunc(A, er=-1.0)
    return U, S, Vh


def contract_svd(T1, T2, indx, er=0.0):
    """``contract_svd(T1, T2, (i1, i2); er)`` (src/svd.jl:7-38); Tensors in, Tensor out."""
    from .tensor_network import Tensor
    if not er >= 0:
        raise ValueError("Error must be positive")
    a = as_c128(T1.data if isinstance(T1, Tensor) else T1)
    b = as_c128(T2.data if isinstance(T2, Tensor) else T2)
    i1, i2 = indx
    d1 = a.shape[i1 - 1] if i1 <= a.ndim else 1  # Julia: size(A, d) == 1 beyond ndims
    d2 = b.shape[i2 - 1] if i2 <= b.ndim else 1
    if d1 != d2:
        raise ValueError("Dimensions of contraction legs do not match")
    _lib.require_device()
    newdim = a.shape[:i1 - 1] + a.shape[i1:] + b.shape[:i2 - 1] + b.shape[i2:]
    out = np.zeros(newdim, dtype=np.complex128, order="F")
    check(lib.qtn_contract_svd(a.ctypes.data_as(C.c_void_p), a.ndim, arr_i64(a.shape), i1,
                               b.ctypes.data_as(C.c_void_p), b.ndim, arr_i64(b.shape), i2, float(er),
                               out.ctypes.data_as(C.c_void_p)))
    return Tensor(out)


def orth_columns(A):
    """EXTENSION: orthonormal basis Q of the columns of a tall matrix (the gauge step of the MPO x MPS
    compression, where the reference would keep ``U`` of ``svd``).  Returns (Q, method) with method
    1 = blocked CholeskyQR2, 2 = Jacobi SVD fallback."""
    _lib.require_device()
    A = as_c128(A)
    if A.ndim != 2 or A.shape[0] < A.shape[1] or A.shape[1] < 1:
        raise ValueError("orth_columns expects an m x n matrix with m >= n >= 1")
    m, n = A.shape
    Q = np.zeros((m, n), dtype=np.complex128, order="F")
    method = C.c_int32(0)
    check(lib.qtn_orth_columns(A.ctypes.data_as(C.c_void_p), m, n, Q.ctypes.data_as(C.c_void_p), C.byref(method)))
    return Q, int(method.value)


def operator_chain(m, M, entry="qtn_mpo_from_matrix"):
    """The SVD chain of ``MPO(m)`` (src/mpo.jl:40-75) / ``decompose!`` (src/decompose.jl:17-48) in ONE library
    call: reshape, permutedims and the left-to-right un-truncated SVDs run on the device, the running
    ``diagm(S) * V'`` never returns to the host.  Returns the M site arrays: (2, 2, b1), (b_i, 2, 2, b_{i+1}), ...,
    (b_{M-1}, 2, 2)."""
    _lib.require_device()
    m = as_c128(m)
    caps, b = [], 1
    for i in range(1, M):
        b = min(4 * b, 4 ** (M - i))
        caps.append(b)
    shapes = [(2, 2, caps[0])] + [(caps[i - 1], 2, 2, caps[i]) for i in range(1, M - 1)] + [(caps[-1], 2, 2)]
    bufs = [np.zeros(int(np.prod(sh)), dtype=np.complex128) for sh in shapes]
    ptrs = (C.c_void_p * M)(*[x.ctypes.data for x in bufs])
    bonds = (C.c_int64 * max(M - 1, 1))()
    check(getattr(lib, entry)(m.ctypes.data_as(C.c_void_p), M, ptrs, bonds))
    bd = [int(bonds[i]) for i in range(M - 1)]
    assert bd == caps, (bd, caps)
    return [np.reshape(x, sh, order="F") for x, sh in zip(bufs, shapes)]
