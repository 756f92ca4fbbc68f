"""``MPO``, ``extend_MPO``, ``apply_MPO`` (src/mpo.jl): symbolic network growth on the
host; the operator-splitting SVDs run on the GPU."""
import numpy as np

from .gates import CircuitGate
from .svd import operator_chain
from .tensor_network import GeneralTensorNetwork, Summation, Tensor, TensorNetwork, shift_pair, shift_summation


class MPO(TensorNetwork):
    def __init__(self, m, contractions=None, openidx=None):
        if contractions is not None:  # src/mpo.jl:16-19
            raise ValueError("Direct conversion to MPS form is not support, please construct MPO from matrix or CircuitGate objects")
        if isinstance(m, CircuitGate):
            m = m.matrix
        m = np.asarray(m)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise AssertionError("MPO needs a square matrix")
        M = m.shape[0].bit_length() - 1
        if not M >= 1:
            raise ValueError("Need at least one qubit to act on.")
        self.tensors, self.contractions = [], []
        if M == 1:
            self.tensors.append(Tensor(np.reshape(m, (2, 2), order="F")))
            self.openidx = [(1, 2), (1, 1)]
            return
        # src/mpo.jl:40-70: the whole reshape / permutedims / sequential-SVD chain is one device call
        for i, site in enumerate(operator_chain(np.reshape(m, (2 ** M, 2 ** M), order="F"), M), 1):
            self.tensors.append(Tensor(site))
            if i < M:
                self.contractions.append(Summation([(i, 3 if i == 1 else 4), (i + 1, 1)]))
        self.openidx = [(M - i + 1, 2) for i in range(1, M)] + [(1, 1)] + [(M - i + 1, 3) for i in range(1, M)] + [(1, 2)]

    def isapprox(self, other):
        return (all(a.isapprox(b) for a, b in zip(self.tensors, other.tensors)) and
                self.contractions == other.contractions and self.openidx == other.openidx)


def _check_wires(iwire, need_sorted):
    if len(set(iwire)) != len(iwire):
        raise ValueError("Repeated wires are not valid.")
    if need_sorted and list(iwire) != sorted(iwire, reverse=True):
        raise ValueError("Wires not sorted")
    if not all(w > 0 for w in iwire):
        raise ValueError("Wires must be positive integers.")


def extend_MPO(mpo, iwire):  # src/mpo.jl:122-173 (mutates `mpo` like the reference)
    iwire = tuple(iwire)
    if not isinstance(mpo, MPO):
        _check_wires(iwire, False)
        if list(iwire) != sorted(iwire, reverse=True):
            raise ValueError("Wires not sorted")
        mpo = MPO(mpo)
    _check_wires(iwire, True)
    M = len(iwire)
    iw = iwire[::-1]
    N = iw[-1] - iw[0] + 1
    if len(mpo.tensors) != M:
        raise AssertionError("MPO length does not match the wires")
    if M == N:
        raise ValueError("MPO is already decomposed in N tensors")
    qwire = list(range(iw[0], iw[-1] + 1))
    pipes = sorted(w for w in qwire if w not in iw)
    qwire.reverse()
    for w in pipes:
        ind = qwire.index(w) + 1
        bond = mpo.tensors[ind - 2].size()[-1]
        pipe = np.reshape(np.kron(np.eye(bond), np.eye(2)), (bond, 2, bond, 2), order="F")
        mpo.tensors.insert(ind - 1, Tensor(np.transpose(pipe, (0, 1, 3, 2)).astype(np.complex128)))
    for i in range(M, N):
        mpo.contractions.append(Summation([(i, 4), (i + 1, 1)]))
        mpo.openidx.insert(0, (i + 1, 2))
        mpo.openidx.insert(i + 1, (i + 1, 3))
    return mpo


def apply_MPO(psi, op, iwire=None):  # src/mpo.jl:184-252
    if isinstance(op, CircuitGate):
        return apply_MPO(psi, op.matrix, op.iwire)
    iwire = tuple(int(w) for w in iwire)
    M = len(iwire)
    if not isinstance(op, MPO):
        _check_wires(iwire, False)
        m = np.asarray(op)
        srt = sorted(iwire)
        if srt != list(iwire):
            order = sorted(range(M), key=lambda a: iwire[a])
            m = np.reshape(m, (2,) * (2 * M), order="F")
            m = np.reshape(np.transpose(m, order + [o + M for o in order]), (2 ** M, 2 ** M), order="F")
        return apply_MPO(psi, MPO(m), tuple(srt))
    if len(set(iwire)) != len(iwire):
        raise ValueError("Repeated wires are not valid.")
    n = len(psi.openidx)
    if not all(0 < w <= n for w in iwire):
        raise ValueError("Wires must be integers between 1 and n (total number of qudits).")
    step = len(psi.tensors)
    rev = iwire[::-1]
    out = GeneralTensorNetwork(list(psi.tensors) + list(op.tensors),
                               list(psi.contractions) + [shift_summation(c, step) for c in op.contractions],
                               list(psi.openidx))
    for i, w in enumerate(rev, 1):
        out.contractions.append(Summation([psi.openidx[w - 1], shift_pair(op.openidx[i + M - 1], step)]))
    for i, w in enumerate(rev, 1):
        out.openidx[w - 1] = shift_pair(op.openidx[i - 1], step)
    return out
