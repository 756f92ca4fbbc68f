"""Oracle restatement of the Qaintessent / Qaintmodels pieces the path needs (test infrastructure).

Qaintessent 0.1.1 and Qaintmodels 0.1.0 are un-vendored dependencies of the
reference (``Manifest.toml:346-358``).  Restated from their published
behaviour and pinned by independent dense linear algebra in the tests
(``kron``-built ``U*psi``, DFT matrix for ``qft_circuit``):

* a circuit gate is ``(iwire, matrix)``; ``iwire[0]`` is the least-significant
  bit of the gate-matrix index; for controlled gates ``iwire = (targets...,
  controls...)`` and the matrix is ``blockdiag(I, U)`` (controls = most
  significant bits) -- the literal CNOT of ``test/test_mpo.jl:81`` pins this;
* wire 1 is the fastest-varying bit of the state vector
  (``src/contract.jl:51-52``, ``src/mps.jl:60-62``).
"""
import numpy as np

X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
S = np.diag([1, 1j]).astype(np.complex128)
T = np.diag([1, np.exp(1j * np.pi / 4)]).astype(np.complex128)
SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def phase_shift(phi):
    return np.diag([1, np.exp(1j * phi)]).astype(np.complex128)


def rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def ry(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def rz(t):
    return np.diag([np.exp(-1j * t / 2), np.exp(1j * t / 2)]).astype(np.complex128)


def controlled(U, ncontrol=1):
    """``matrix(ControlledGate(U, ncontrol))``: U acts iff all (most significant) control bits are 1."""
    U = np.asarray(U, dtype=np.complex128)
    n = U.shape[0] * 2 ** ncontrol
    CU = np.eye(n, dtype=np.complex128)
    CU[n - U.shape[0]:, n - U.shape[0]:] = U
    return CU


class CircuitGate:
    def __init__(self, iwire, matrix):
        self.iwire = tuple(int(w) for w in (iwire if isinstance(iwire, (tuple, list)) else (iwire,)))
        self.matrix = np.asarray(matrix, dtype=np.complex128)
        assert len(set(self.iwire)) == len(self.iwire)
        assert self.matrix.shape == (2 ** len(self.iwire),) * 2

    @property
    def M(self):
        return len(self.iwire)

    def req_wires(self):
        return max(self.iwire)


def circuit_gate(target, U, control=()):
    """``circuit_gate(target, U[, control])``: iwire = (targets..., controls...)."""
    t = tuple(target) if isinstance(target, (tuple, list)) else (target,)
    c = tuple(control) if isinstance(control, (tuple, list)) else (control,)
    U = np.asarray(U, dtype=np.complex128)
    return CircuitGate(t + c, controlled(U, len(c)) if c else U)


def qft_circuit(N):
    """``Qaintmodels.qft_circuit(N)``: H(i), controlled phase 2pi/2^(j-i+1) (target i, control j), swaps."""
    cgc = []
    for i in range(1, N + 1):
        cgc.append(circuit_gate(i, H))
        for j in range(i + 1, N + 1):
            cgc.append(circuit_gate(i, phase_shift(2 * np.pi / 2 ** (j - i + 1)), j))
    for i in range(1, N // 2 + 1):
        cgc.append(circuit_gate((i, N - i + 1), SWAP))
    return cgc


def apply(psi, cgc):
    """``Qaintessent.apply(psi, cgc)``: dense state-vector simulator, wire 1 fastest."""
    psi = np.asarray(psi, dtype=np.complex128)
    N = int(round(np.log2(psi.size)))
    assert 2 ** N == psi.size
    if isinstance(cgc, CircuitGate):
        cgc = [cgc]
    t = np.reshape(psi, (2,) * N, order="F")  # axis w-1 <-> wire w
    for cg in cgc:
        M = cg.M
        G = np.reshape(cg.matrix, (2,) * (2 * M), order="F")  # (row bits lsb..msb, col bits lsb..msb)
        axes = [w - 1 for w in cg.iwire]
        out = np.tensordot(G, t, axes=(list(range(M, 2 * M)), axes))  # (row bits..., rest...)
        rest = [a for a in range(N) if a not in axes]
        cur = axes + rest  # current axis k holds original axis cur[k]
        t = np.transpose(out, [cur.index(a) for a in range(N)])
    return np.reshape(t, (-1,), order="F")


def interaction_graph(cgc):
    """src/network2graph.jl:161-174."""
    from .lightgraphs import Graph
    N = max(cg.req_wires() for cg in cgc)
    G = Graph(N)
    for cg in cgc:
        for j, i1 in enumerate(cg.iwire):
            for i2 in cg.iwire[:j]:
                G.add_edge(i1, i2)
    return G
