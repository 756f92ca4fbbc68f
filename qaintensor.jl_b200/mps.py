"""``MPS`` and friends (src/mps.jl, src/switch.jl).  Every SVD goes through the library's
Jacobi SVD (``qtn_svd_trunc``), every two-site contraction through ``qtn_contract_svd``."""
import numpy as np

from .svd import contract_svd, svd
from .tensor_network import Summation, Tensor, TensorNetwork, is_power_two

_E_FIRST = "Tensor objects first leg must contract with last leg of previous Tensor object"
_E_LAST = "Tensor objects last leg must contract with first leg of next Tensor object"
_E_LEGS = "Each Tensor object in MPS form can only have 2 or 3 legs"


def _check(tensors, contractions):
    for c in contractions:
        if len(c.idx) != 2:
            raise AssertionError("MPS contractions join exactly two legs")
        if c.idx[0][1] != tensors[c.idx[0][0] - 1].ndims():
            raise ValueError(_E_FIRST)
        if c.idx[1][1] != 1:
            raise ValueError(_E_LAST)


class MPS(TensorNetwork):
    """``MPS <: TensorNetwork`` with the validating constructor of src/mps.jl:7-29."""

    def __init__(self, tensors, contractions=None, openidx=None):
        if contractions is None:  # MPS(psi::Vector{ComplexF64})
            built = _from_vector(tensors)
            tensors, contractions, openidx = built
        _check(tensors, contractions)
        for t in tensors:
            if t.ndims() not in (2, 3):
                raise ValueError(_E_LEGS)
        self.tensors = list(tensors)
        self.contractions = list(contractions)
        self.openidx = [(int(a), int(b)) for (a, b) in openidx]

    def copy(self):  # shallow (src/mps.jl:203)
        return MPS(list(self.tensors), list(self.contractions), list(self.openidx))


def check_mps(mps):  # src/mps.jl:37-48
    for t in mps.tensors:
        if t.ndims() not in (2, 3):
            raise ValueError(_E_LEGS)
    _check(mps.tensors, mps.contractions)


def _from_vector(psi):  # src/mps.jl:55-89: sequential thin SVD, no truncation
    psi = np.asarray(psi, dtype=np.complex128).reshape(-1)
    if not is_power_two(psi.size):
        raise ValueError("Input state must have length 2^N")
    M = psi.size.bit_length() - 1
    if M >= 2:
        return _from_vector_device(psi, M)
    tensors, contractions, openidx = [], [], [(1, 1)]
    U, S, Vh = svd(np.reshape(psi, (2, -1), order="F"))
    tensors.append(Tensor(U))
    rest = S[:, None] * Vh
    lbond, lastleg = len(S), 2
    for bit in range(2, M):
        U, S, Vh = svd(np.reshape(rest, (lbond * 2, -1), order="F"))
        tensors.append(Tensor(np.reshape(U, (lbond, 2, len(S)), order="F")))
        contractions.append(Summation([(bit - 1, lastleg), (bit, 1)]))
        openidx.append((bit, 2))
        rest = S[:, None] * Vh
        lbond, lastleg = len(S), 3
    tensors.append(Tensor(rest))
    contractions.append(Summation([(M - 1, lastleg), (M, 1)]))
    openidx.append((M, 2))
    return tensors, contractions, openidx


def _from_vector_device(psi, M):
    """The whole SVD chain of ``MPS(psi)`` in one library call (``qtn_mps_from_vector``)."""
    import ctypes as C
    from . import _lib
    _lib.require_device()
    bonds_cap = [min(2 ** i, 2 ** (M - i)) for i in range(1, M)]
    shapes = [(2, bonds_cap[0])] + [(bonds_cap[i - 1], 2, bonds_cap[i]) for i in range(1, M - 1)] + [(bonds_cap[-1], 2)]
    bufs = [np.zeros(int(np.prod(s)), dtype=np.complex128) for s in shapes]
    ptrs = (C.c_void_p * M)(*[b.ctypes.data for b in bufs])
    bonds = (C.c_int64 * max(M - 1, 1))()
    _lib.check(_lib.lib.qtn_mps_from_vector(_lib.as_c128(psi).ctypes.data_as(C.c_void_p), M, ptrs, bonds))
    b = [int(bonds[i]) for i in range(M - 1)]
    tensors = []
    for i in range(M):
        shp = (2, b[0]) if i == 0 else ((b[-1], 2) if i == M - 1 else (b[i - 1], 2, b[i]))
        tensors.append(Tensor(np.reshape(bufs[i][:int(np.prod(shp))], shp, order="F")))
    contractions = [Summation([(1, 2), (2, 1)])] + [Summation([(i, 3), (i + 1, 1)]) for i in range(2, M)]
    openidx = [(1, 1)] + [(i, 2) for i in range(2, M + 1)]
    return tensors, contractions, openidx


def OpenMPS(T, N=None):  # src/mps.jl:99-121
    T = [T] * N if N is not None else list(T)
    if any(t.ndims() != 3 for t in T):
        raise ValueError("Tensors must have 3 legs")
    n = len(T)
    return MPS(T, [Summation([(i, 3), (i + 1, 1)]) for i in range(1, n)],
               [(1, 1)] + [(i, 2) for i in range(1, n + 1)] + [(n, 3)])


def ClosedMPS(T, Tmiddle=None, Tend=None, N=None):  # src/mps.jl:130-155
    T = [T] + [Tmiddle] * (N - 2) + [Tend] if Tmiddle is not None else list(T)
    n = len(T)
    if T[0].ndims() != 2:
        raise ValueError("First tensor must have 2 legs")
    if any(t.ndims() != 3 for t in T[1:-1]):
        raise ValueError("Tensors must have 3 legs, except the first and last one")
    if T[-1].ndims() != 2:
        raise ValueError("Last tensor must have 2 legs")
    cons = [Summation([(1, 2), (2, 1)])] + [Summation([(i, 3), (i + 1, 1)]) for i in range(2, n)]
    return MPS(T, cons, [(1, 1)] + [(i, 2) for i in range(2, n + 1)])


def PeriodicMPS(T, N=None):  # src/mps.jl:163-182
    T = [T] * N if N is not None else list(T)
    if any(t.ndims() != 3 for t in T):
        raise AssertionError("PeriodicMPS tensors must have 3 legs")
    n = len(T)
    cons = [Summation([(i, 3), (i + 1, 1)]) for i in range(1, n)] + [Summation([(n, 3), (1, 1)])]
    return MPS(T, cons, [(i, 2) for i in range(1, n + 1)])


def contract_svd_mps(tn, er=0.0):  # src/mps.jl:190-201
    if not er >= 0:
        raise ValueError("Error must be positive")
    n = len(tn.tensors)
    periodic = (Summation([(n, 3), (1, 1)]), Summation([(1, 1), (n, 3)]))
    if any(s in periodic for s in tn.contractions):
        raise ValueError("Function doesn't support periodic boundary conditions for now")
    # the fold tcontract = contract_svd(tcontract, T_j, (ndims, 1); er) in one library call (qtn_contract_svd_fold):
    # the running tensor stays on the device
    import ctypes as C
    from . import _lib
    _lib.require_device()
    arrs = [_lib.as_c128(t.data) for t in tn.tensors]
    for a, b in zip(arrs[:-1], arrs[1:]):
        if a.shape[-1] != b.shape[0]:
            raise ValueError("Dimensions of contraction legs do not match")
    shape = arrs[0].shape[:-1] if n > 1 else arrs[0].shape
    for j in range(1, n):
        shape = shape + (arrs[j].shape[1:-1] if j < n - 1 else arrs[j].shape[1:])
    out = np.zeros(shape, dtype=np.complex128, order="F")
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    _lib.check(_lib.lib.qtn_contract_svd_fold(n, ptrs, _lib.arr_i64([a.size for a in arrs]), _lib.arr_i64([a.shape[0] for a in arrs]),
                                              _lib.arr_i64([a.shape[-1] for a in arrs]), float(er),
                                              out.ctypes.data_as(C.c_void_p), out.size))
    return out


# ---- src/switch.jl ---------------------------------------------------------------------
def _switch_adjacent(mps, i):  # switch!(mps, i) src/switch.jl:18-56
    from .contract import permutedims
    if i < 1 or i + 1 > len(mps.tensors):
        raise IndexError("BoundsError: attempt to access %d-element MPS at index %d" % (len(mps.tensors), i + 1))
    T1, T2 = mps.tensors[i - 1], mps.tensors[i]
    d1, d2 = T1.size(), T2.size()
    if T1.ndims() == 2 and T2.ndims() == 2:   # the reference's permutedims(T, [2,1,3]) of a 2-leg T throws as well
        raise ValueError("switch! of a two-tensor MPS: the contracted pair has no third leg to permute")
    # src/switch.jl:26-52 in one library call (contract_svd, leg exchange, svd, diagm(S) * V' on the device)
    import ctypes as C
    from . import _lib
    _lib.require_device()
    a, b = _lib.as_c128(T1.data), _lib.as_c128(T2.data)
    if d1[-1] != d2[0]:
        raise ValueError("Dimensions of contraction legs do not match")
    l1 = d1[0] if T1.ndims() == 3 else 0
    r2 = d2[-1] if T2.ndims() == 3 else 0
    bond = min(2 * max(l1, 1), 2 * max(r2, 1))
    U = np.zeros((2, bond) if l1 == 0 else (l1, 2, bond), dtype=np.complex128, order="F")
    V = np.zeros((bond, 2) if r2 == 0 else (bond, 2, r2), dtype=np.complex128, order="F")
    kb = C.c_int64(0)
    _lib.check(_lib.lib.qtn_mps_switch_adjacent(a.ctypes.data_as(C.c_void_p), l1, d1[-1], b.ctypes.data_as(C.c_void_p), r2,
                                                U.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), C.byref(kb)))
    assert kb.value == bond
    mps.tensors[i - 1] = Tensor(U)
    mps.tensors[i] = Tensor(V)


def switch(mps, i, j=None):
    """``switch!``: adjacent swap (one index), wire swap (two), or a list of (i, j) tuples."""
    if isinstance(i, (list, tuple)) and j is None:
        check_mps(mps)
        for (a, b) in i:
            switch(mps, a, b)
        return
    if j is None:
        return _switch_adjacent(mps, i)
    check_mps(mps)
    if not (i > 0 and j > 0):
        raise ValueError("Wire indices `i` and `j` must be positive")
    n = len(mps.tensors)
    if not (i <= n and j <= n):
        raise ValueError("Indices to swap `i` and `j` must be less than or equal to the number of open wires in MPS")
    if i == j:
        return
    lo, hi = sorted((n - i + 1, n - j + 1))
    for a in range(lo, hi):
        _switch_adjacent(mps, a)
    for b in range(hi - 2, lo - 1, -1):
        _switch_adjacent(mps, b)


def permute(mps, order):  # Base.permute!(mps, order) src/switch.jl:94-106
    check_mps(mps)
    n = len(mps.tensors)
    if len(order) != n:
        raise ValueError("Given permutation must be same length as number of Tensors in MPS")
    if len(set(order)) != len(order):
        raise ValueError("Permutation order cannot contain repeat values")
    if not all(x > 0 for x in order):
        raise ValueError("Permutation order can only contain positive values")
    if max(order) != n:
        raise ValueError("Wire numbers in permutation order cannot exceed number of wires in MPS")
    seq = list(range(1, n + 1))
    for i in range(1, n + 1):
        loc = seq.index(order[i - 1]) + 1
        switch(mps, i, loc)
        seq[loc - 1], seq[i - 1] = seq[i - 1], seq[loc - 1]
