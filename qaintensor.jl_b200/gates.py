"""Gate set and circuit-gate container the path needs (host-side; mirrors the names
re-exported from Qaintessent / Qaintmodels at src/Qaintensor.jl:9-46).

Convention (pinned by the literal CNOT of test/test_mpo.jl:81): ``iwire[0]`` is the
least-significant bit of the gate-matrix index; controlled gates list
``(targets..., controls...)`` and their matrix is ``blockdiag(1, U)``.
"""
import math

import numpy as np

_c = np.complex128
X = np.array([[0, 1], [1, 0]], dtype=_c)
Y = np.array([[0, -1j], [1j, 0]], dtype=_c)
Z = np.array([[1, 0], [0, -1]], dtype=_c)
HadamardGate = np.array([[1, 1], [1, -1]], dtype=_c) / math.sqrt(2.0)
SGate = np.diag([1, 1j]).astype(_c)
TGate = np.diag([1, np.exp(0.25j * math.pi)]).astype(_c)
SdagGate = SGate.conj().T
TdagGate = TGate.conj().T
SwapGate = np.eye(4, dtype=_c)[[0, 2, 1, 3]]


def PhaseShiftGate(phi):
    return np.diag([1, np.exp(1j * phi)]).astype(_c)


def RxGate(theta):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=_c)


def RyGate(theta):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -s], [s, c]], dtype=_c)


def RzGate(theta):
    return np.diag([np.exp(-0.5j * theta), np.exp(0.5j * theta)]).astype(_c)


def ControlledGate(U, ncontrol=1):
    U = np.asarray(U, dtype=_c)
    n = U.shape[0] << ncontrol
    out = np.eye(n, dtype=_c)
    out[n - U.shape[0]:, n - U.shape[0]:] = U
    return out


controlled_not = lambda: ControlledGate(X)  # noqa: E731


class CircuitGate:
    def __init__(self, iwire, matrix):
        if not isinstance(iwire, (tuple, list)):
            iwire = (iwire,)
        self.iwire = tuple(int(w) for w in iwire)
        self.matrix = np.asarray(matrix, dtype=_c)
        if len(set(self.iwire)) != len(self.iwire):
            raise ValueError("Repeated wires are not valid.")
        if self.matrix.shape != (1 << len(self.iwire),) * 2:
            raise ValueError("gate matrix does not match the number of wires")

    def req_wires(self):
        return max(self.iwire)


def circuit_gate(target, U, control=()):
    t = tuple(target) if isinstance(target, (tuple, list)) else (target,)
    c = tuple(control) if isinstance(control, (tuple, list)) else (control,)
    return CircuitGate(t + c, ControlledGate(U, len(c)) if c else U)


def qft_circuit(N):
    """``Qaintmodels.qft_circuit``: H(i); controlled phase 2*pi/2^(j-i+1) with target i,
    control j (j > i); finally Swap(i, N-i+1)."""
    out = []
    for i in range(1, N + 1):
        out.append(circuit_gate(i, HadamardGate))
        for j in range(i + 1, N + 1):
            out.append(circuit_gate(i, PhaseShiftGate(2.0 * math.pi / (1 << (j - i + 1))), j))
    for i in range(1, N // 2 + 1):
        out.append(circuit_gate((i, N - i + 1), SwapGate))
    return out
