"""Oracle restatement of ``src/mps.jl`` and ``src/switch.jl`` (test infrastructure)."""
import numpy as np

from .network import Network, Summation, Tensor, is_power_two
from .svd import contract_svd, svd


class MPS(Network):
    """``MPS <: TensorNetwork`` with the validating inner constructor (src/mps.jl:7-29)."""

    def __init__(self, tensors, contractions, openidx):
        for c in contractions:
            assert len(c.idx) == 2
            if not c.idx[0][1] == tensors[c.idx[0][0] - 1].ndims:
                raise ValueError("Tensor objects first leg must contract with last leg of previous Tensor object")
            if not c.idx[1][1] == 1:
                raise ValueError("Tensor objects last leg must contract with first leg of next Tensor object")
        for t in tensors:
            # src/mps.jl:24 relies on operator precedence but accepts exactly ranks 2 and 3 (quirk Q5)
            if t.ndims not in (2, 3):
                raise ValueError("Each Tensor object in MPS form can only have 2 or 3 legs")
        super().__init__(tensors, contractions, openidx)


def check_mps(mps):  # src/mps.jl:37-48
    for t in mps.tensors:
        if t.ndims not in (2, 3):
            raise ValueError("Each Tensor object in MPS form can only have 2 or 3 legs")
    for c in mps.contractions:
        assert len(c.idx) == 2
        if not c.idx[0][1] == mps.tensors[c.idx[0][0] - 1].ndims:
            raise ValueError("Tensor objects first leg must contract with last leg of previous Tensor object")
        if not c.idx[1][1] == 1:
            raise ValueError("Tensor objects last leg must contract with first leg of next Tensor object")


def mps_from_vector(psi):  # ``MPS(psi::Vector{ComplexF64})`` src/mps.jl:55-89
    psi = np.asarray(psi, dtype=np.complex128)
    if not is_power_two(psi.size):
        raise ValueError("Input state must have length 2^N")
    M = int(round(np.log2(psi.size)))
    tensors, contractions, openidx = [], [], []
    m = np.reshape(psi, (2, 2 ** (M - 1)), order="F")
    U, S, Vh = svd(m)
    tensors.append(Tensor(U))
    openidx.append((1, 1))
    lbond = len(S)
    m = S[:, None] * Vh
    lastbit = 2
    for bit in range(2, M):
        m = np.reshape(m, (lbond * 2, 2 ** (M - bit)), order="F")
        U, S, Vh = svd(m)
        rbond = len(S)
        m = S[:, None] * Vh
        tensors.append(Tensor(np.reshape(U, (lbond, 2, rbond), order="F")))
        contractions.append(Summation([(bit - 1, lastbit), (bit, 1)]))
        openidx.append((bit, 2))
        lbond = rbond
        lastbit = 3
    tensors.append(Tensor(m))
    contractions.append(Summation([(M - 1, lastbit), (M, 1)]))
    openidx.append((M, 2))
    return MPS(tensors, contractions, openidx)


def OpenMPS(T, N=None):  # src/mps.jl:99-121
    if N is not None:
        T = [T] * N
    l = len(T)
    for t in T:
        if t.ndims != 3:
            raise ValueError("Tensors must have 3 legs")
    contractions = [Summation([(i, 3), (i + 1, 1)]) for i in range(1, l)]
    openidx = [(1, 1)] + [(i, 2) for i in range(1, l + 1)] + [(l, 3)]
    return MPS(T, contractions, openidx)


def ClosedMPS(T, Tmiddle=None, Tend=None, N=None):  # src/mps.jl:130-155
    if Tmiddle is not None:
        T = [T] + [Tmiddle] * (N - 2) + [Tend]
    l = len(T)
    if T[0].ndims != 2:
        raise ValueError("First tensor must have 2 legs")
    for i in range(1, l - 1):
        if T[i].ndims != 3:
            raise ValueError("Tensors must have 3 legs, except the first and last one")
    if T[l - 1].ndims != 2:
        raise ValueError("Last tensor must have 2 legs")
    contractions = [Summation([(1, 2), (2, 1)])] + [Summation([(i, 3), (i + 1, 1)]) for i in range(2, l)]
    openidx = [(1, 1)] + [(i, 2) for i in range(2, l + 1)]
    return MPS(T, contractions, openidx)


def PeriodicMPS(T, N=None):  # src/mps.jl:163-182
    if N is not None:
        T = [T] * N
    l = len(T)
    for t in T:
        assert t.ndims == 3
    contractions = [Summation([(i, 3), (i + 1, 1)]) for i in range(1, l)] + [Summation([(l, 3), (1, 1)])]
    openidx = [(i, 2) for i in range(1, l + 1)]
    return MPS(T, contractions, openidx)


def contract_svd_mps(tn, er=0.0):  # src/mps.jl:190-201
    if not er >= 0:
        raise ValueError("Error must be positive")
    l = len(tn.tensors)
    if any(s == Summation([(l, 3), (1, 1)]) for s in tn.contractions) or \
       any(s == Summation([(1, 1), (l, 3)]) for s in tn.contractions):
        raise ValueError("Function doesn't support periodic boundary conditions for now")
    acc = tn.tensors[0].data
    for j in range(1, l):
        acc = contract_svd(acc, tn.tensors[j].data, (acc.ndim, 1), er=er)
    return acc


# ----------------------------------------------------------------------------
# src/switch.jl
# ----------------------------------------------------------------------------
def switch_adjacent(mps, i):  # ``switch!(mps, i)`` src/switch.jl:18-56
    T1 = mps.tensors[i - 1]  # IndexError mirrors the reference's BoundsError (quirk Q5)
    if i < 1 or i + 1 > len(mps.tensors):
        raise IndexError("BoundsError")
    T2 = mps.tensors[i]
    d1, d2 = T1.size, T2.size
    T = contract_svd(T1.data, T2.data, (T1.ndims, 1))
    if T1.ndims == 2:
        T = np.reshape(np.transpose(T, (1, 0, 2)), (2, 2 * d2[-1]), order="F")
    elif T2.ndims == 2:
        T = np.reshape(np.transpose(T, (0, 2, 1)), (2 * d1[0], 2), order="F")
    else:
        T = np.reshape(np.transpose(T, (0, 2, 1, 3)), (2 * d1[0], 2 * d2[-1]), order="F")
    U, S, Vh = svd(T)
    bond = len(S)
    V = S[:, None] * Vh
    if T1.ndims == 2:
        U = np.reshape(U, (2, bond), order="F")
        V = np.reshape(V, (bond, 2, d2[-1]), order="F")
    elif T2.ndims == 2:
        U = np.reshape(U, (d1[0], 2, bond), order="F")
    else:
        U = np.reshape(U, (d1[0], 2, bond), order="F")
        V = np.reshape(V, (bond, 2, d2[-1]), order="F")
    mps.tensors[i - 1] = Tensor(U)
    mps.tensors[i] = Tensor(V)


def switch(mps, i, j):  # ``switch!(mps, i, j)`` src/switch.jl:63-86
    check_mps(mps)
    if not (i > 0 and j > 0):
        raise ValueError("Wire indices `i` and `j` must be positive")
    n = len(mps.tensors)
    if not ((i <= n) and (j <= n)):
        raise ValueError("Indices to swap `i` and `j` must be less than or equal to the number of open wires in MPS")
    if i == j:
        return
    i = n - i + 1
    j = n - j + 1
    lo, hi = (i, j) if i < j else (j, i)
    for a in range(lo, hi):
        switch_adjacent(mps, a)
    for b in range(hi - 2, lo - 1, -1):
        switch_adjacent(mps, b)


def permute(mps, order):  # ``Base.permute!(mps, order)`` src/switch.jl:94-106
    check_mps(mps)
    n = len(mps.tensors)
    if len(order) != n:
        raise ValueError("Given permutation must be same length as number of Tensors in MPS")
    if len(set(order)) != len(order):
        raise ValueError("Permutation order cannot contain repeat values")
    if not all(x > 0 for x in order):
        raise ValueError("Permutation order can only contain positive values")
    if max(order) != len(order):
        raise ValueError("Wire numbers in permutation order cannot exceed number of wires in MPS")
    seq = list(range(1, n + 1))
    for i in range(1, n + 1):
        loc = seq.index(order[i - 1]) + 1
        switch(mps, i, loc)
        seq[loc - 1], seq[i - 1] = seq[i - 1], seq[loc - 1]
