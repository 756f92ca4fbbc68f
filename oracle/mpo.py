"""Oracle restatement of ``src/mpo.jl``, ``src/decompose.jl`` and ``src/tensor_circuit.jl`` (test infrastructure)."""
import numpy as np

from .gates import CircuitGate
from .network import Network, Summation, Tensor, shift_pair, shift_summation
from .svd import svd


class MPO(Network):
    """``MPO(m::AbstractMatrix)`` (src/mpo.jl:27-90).  The (tensors, contractions,
    openidx) constructor of the reference always errors (src/mpo.jl:16-19)."""

    def __init__(self, m):
        m = np.asarray(m)
        assert m.shape[0] == m.shape[1]
        M = int(round(np.log2(m.shape[0])))
        if not M >= 1:
            raise ValueError("Need at least one qubit to act on.")
        t, con, openidx = [], [], []
        if M > 1:
            m = np.reshape(m, (2,) * (2 * M), order="F")
            dims = []
            for i in range(1, M + 1):
                dims += [i, i + M]
            m = np.transpose(m, [d - 1 for d in dims])
            bond = 1
            m = np.reshape(m, (4 * bond, -1), order="F")
            U, S, Vh = svd(m)
            bond = len(S)
            m = S[:, None] * Vh
            t.append(Tensor(np.reshape(U, (2, 2, bond), order="F")))
            con.append(Summation([(1, 3), (2, 1)]))
            for i in range(2, M):
                m = np.reshape(m, (bond * 4, -1), order="F")
                U, S, Vh = svd(m)
                m = S[:, None] * Vh
                nb = len(S)
                t.append(Tensor(np.reshape(U, (bond, 2, 2, nb), order="F")))
                con.append(Summation([(i, 4), (i + 1, 1)]))
                bond = nb
            t.append(Tensor(np.reshape(m, (bond, 2, 2), order="F")))
            for i in range(1, M):
                openidx.append((M - i + 1, 2))
            openidx.append((1, 1))
            for i in range(1, M):
                openidx.append((M - i + 1, 3))
            openidx.append((1, 2))
        else:
            t.append(Tensor(np.reshape(m, (2, 2), order="F")))
            openidx = [(1, 2), (1, 1)]
        super().__init__(t, con, openidx)

    def copy(self):
        new = object.__new__(MPO)
        Network.__init__(new, list(self.tensors), list(self.contractions), list(self.openidx))
        return new


def _check_wires_sorted_desc(iwire):
    if len(set(iwire)) != len(iwire):
        raise ValueError("Repeated wires are not valid.")
    if list(iwire) != sorted(iwire, reverse=True):
        raise ValueError("Wires not sorted")
    if not all(w > 0 for w in iwire):
        raise ValueError("Wires must be positive integers.")


def extend_MPO(mpo, iwire):  # src/mpo.jl:122-157 (mutates its argument, quirk Q6)
    if not isinstance(mpo, MPO):  # matrix method src/mpo.jl:168-173
        iw = tuple(iwire)
        if len(set(iw)) != len(iw):
            raise ValueError("Repeated wires are not valid.")
        if not all(w > 0 for w in iw):
            raise ValueError("Wires must be positive integers.")
        if list(iw) != sorted(iw, reverse=True):
            raise ValueError("Wires not sorted")
        mpo = MPO(mpo)
    iwire = tuple(iwire)
    M = len(iwire)
    _check_wires_sorted_desc(iwire)
    iwire = iwire[::-1]
    N = iwire[-1] - iwire[0] + 1
    assert len(mpo.tensors) == M
    if M == N:
        raise ValueError("MPO is already decomposed in N tensors")
    d = 2
    qwire = list(range(iwire[0], iwire[-1] + 1))
    pipeswire = sorted(w for w in qwire if w not in iwire)
    qwire = qwire[::-1]
    for w in pipeswire:
        ind = qwire.index(w) + 1
        bond = mpo.tensors[ind - 2].size[-1]
        Vpipe = np.reshape(np.kron(np.eye(bond), np.eye(d)), (bond, d, bond, d), order="F")
        Vpipe = np.transpose(Vpipe, (0, 1, 3, 2)).astype(np.complex128)
        mpo.tensors.insert(ind - 1, Tensor(Vpipe))
    for i in range(M, N):
        mpo.contractions.append(Summation([(i, 4), (i + 1, 1)]))
        mpo.openidx.insert(0, (i + 1, 2))
        mpo.openidx.insert(i + 1, (i + 1, 3))
    return mpo


def apply_MPO(psi, op, iwire=None):  # src/mpo.jl:184-252
    if isinstance(op, CircuitGate):  # :248-252
        return apply_MPO(psi, op.matrix, op.iwire)
    iwire = tuple(iwire)
    M = len(iwire)
    if not isinstance(op, MPO):  # matrix method :228-241
        if len(set(iwire)) != len(iwire):
            raise ValueError("Repeated wires are not valid.")
        if not all(w > 0 for w in iwire):
            raise ValueError("Wires must be positive integers.")
        m = np.asarray(op)
        iwire_sorted = sorted(iwire)
        if iwire_sorted != list(iwire):
            sort_wires = list(np.argsort(np.array(iwire), kind="stable"))
            perm = sort_wires + [s + M for s in sort_wires]
            m = np.reshape(m, (2,) * (2 * M), order="F")
            m = np.transpose(m, perm)
            m = np.reshape(m, (2 ** M, 2 ** M), order="F")
        return apply_MPO(psi, MPO(m), tuple(iwire_sorted))
    mpo = op
    if len(set(iwire)) != len(iwire):
        raise ValueError("Repeated wires are not valid.")
    n = len(psi.openidx)
    if not all(0 < w <= n for w in iwire):
        raise ValueError("Wires must be integers between 1 and n (total number of qudits).")
    step = len(psi.tensors)
    N = len(iwire)
    iwire = iwire[::-1]
    # `if M < N` (src/mpo.jl:196-202) is dead: N == M always (quirk Q4)
    out = Network(list(psi.tensors) + list(mpo.tensors),
                  list(psi.contractions) + [shift_summation(c, step) for c in mpo.contractions],
                  list(psi.openidx))
    for i, w in enumerate(iwire, 1):
        out.contractions.append(Summation([psi.openidx[w - 1], shift_pair(mpo.openidx[i + N - 1], step)]))
    for i, q in enumerate(iwire, 1):
        out.openidx[q - 1] = shift_pair(mpo.openidx[i - 1], step)
    return out


def decompose(cg):  # ``decompose!(cg)`` src/decompose.jl:6-52
    M = cg.M
    if not M > 1:
        raise ValueError("Only decompose Circuit Gates that apply to multiple wires")
    m = np.array(cg.matrix)
    t, w, c = [], [], []
    m = np.reshape(m, (2,) * (2 * M), order="F")
    dims = []
    for i in range(1, M + 1):
        dims += [i, i + M]
    m = np.transpose(m, [d - 1 for d in dims])
    bond = 1
    m = np.reshape(m, (4 * bond, -1), order="F")
    U, S, Vh = svd(m)
    bond = len(S)
    m = S[:, None] * Vh
    t.append(Tensor(np.reshape(U, (2, 2, bond), order="F")))
    w.append(cg.iwire[0])
    c.append(0)
    c.append(3)
    for i in range(2, M):
        m = np.reshape(m, (bond * 4, -1), order="F")
        U, S, Vh = svd(m)
        m = S[:, None] * Vh
        nb = len(S)
        t.append(Tensor(np.reshape(U, (bond, 2, 2, nb), order="F")))
        w.append(cg.iwire[i - 1])
        c.append(4)
        bond = nb
    t.append(Tensor(np.reshape(m, (bond, 2, 2), order="F")))
    w.append(cg.iwire[M - 1])
    return t, c, w


def tensor_circuit(psi, cgc, is_decompose=False):  # ``tensor_circuit!`` src/tensor_circuit.jl:14-80
    if isinstance(cgc, CircuitGate):
        cgc = [cgc]
    for cg in cgc:
        M = cg.M
        assert cg.req_wires() <= len(psi.openidx)
        if M > 1 and is_decompose:
            ts, cs, ws = decompose(cg)
            for i, (t, c, w) in enumerate(zip(ts, cs, ws), 1):  # zip stops at the shortest (len M)
                psi.tensors.append(t)
                nt = len(psi.tensors)
                psi.contractions.append(Summation([psi.openidx[w - 1], (nt, 2 if i == 1 else 3)]))
                if c > 0:
                    psi.contractions.append(Summation([(nt - 1, c), (nt, 1)]))
                psi.openidx[w - 1] = (nt, 1 if i == 1 else 2)
            continue
        psi.tensors.append(Tensor(np.reshape(cg.matrix, (2,) * (2 * M), order="F")))
        nt = len(psi.tensors)
        for i, w in enumerate(cg.iwire, 1):
            psi.contractions.append(Summation([psi.openidx[w - 1], (nt, i)]))
            psi.openidx[w - 1] = (nt, M + i)
