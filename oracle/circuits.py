"""Oracle: synthetic circuit / network generators (test infrastructure).

EXTENSIONS: the reference has no random-circuit generator.
Networks are assembled with the reference's own ``tensor_circuit!`` restatement
(``src/tensor_circuit.jl:44-51``) on top of rank-1 ket tensors, then closed with
rank-1 bra tensors -- the Markov-Shi form ``optimize_contraction_order!`` is
documented for (``src/network2graph.jl:450-472``).
"""
import numpy as np

from .gates import CircuitGate, qft_circuit
from .mpo import tensor_circuit
from .network import Network, Summation, Tensor


def haar_unitary(n, rng):
    """Haar-random n x n unitary: QR of a complex Ginibre matrix, R's diagonal phases fixed."""
    z = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(2.0)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def brickwork_gates(nq, depth, rng):
    """Layer l (0-based) even: pairs (1,2),(3,4),...; odd: (2,3),(4,5),..."""
    gates = []
    for layer in range(depth):
        for q in range(1 + (layer % 2), nq, 2):
            gates.append(CircuitGate((q, q + 1), haar_unitary(4, rng)))
    return gates


def rqc2d_gates(rows, cols, cycles, rng):
    """Coupler pattern A, B, C, D repeating; qubit (r, c) is wire r*cols + c + 1.

    A: horizontal couplers starting at even columns, B: vertical at even rows,
    C: horizontal at odd columns, D: vertical at odd rows (0-based parity).
    """
    gates = []
    for cyc in range(cycles):
        kind = cyc % 4
        if kind in (0, 2):
            for r in range(rows):
                for c in range(kind // 2, cols - 1, 2):
                    q = r * cols + c + 1
                    gates.append(CircuitGate((q, q + 1), haar_unitary(4, rng)))
        else:
            for r in range(kind // 2, rows - 1, 2):
                for c in range(cols):
                    q = r * cols + c + 1
                    gates.append(CircuitGate((q, q + cols), haar_unitary(4, rng)))
    return gates


def product_input(nq, vectors=None):
    ket0 = np.array([1.0, 0.0], dtype=np.complex128)
    tensors = [Tensor(ket0.copy() if vectors is None else np.asarray(vectors[i], dtype=np.complex128))
               for i in range(nq)]
    return Network(tensors, [], [(i, 1) for i in range(1, nq + 1)])


def close_with_bitstring(net, bits):
    """Project open wire w onto <bits[w-1]| by a rank-1 tensor; openidx becomes empty."""
    for w, b in enumerate(bits, 1):
        v = np.zeros(2, dtype=np.complex128)
        v[int(b)] = 1.0
        net.tensors.append(Tensor(v))
        net.contractions.append(Summation([net.openidx[w - 1], (len(net.tensors), 1)]))
    net.openidx = []
    return net


def amplitude_network(nq, gates, bits):
    net = product_input(nq)
    tensor_circuit(net, gates)
    return close_with_bitstring(net, bits)


def cfg_seed(cfg, index=0):
    return 20261017 + 1000 * cfg + index


def cfg1_qft_network(nq=12, seed=None):
    """Config 1: QFT on a seeded random product state, all wires left open."""
    rng = np.random.default_rng(cfg_seed(1) if seed is None else seed)
    vecs = rng.standard_normal((nq, 2)) + 1j * rng.standard_normal((nq, 2))
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    net = product_input(nq, vecs)
    tensor_circuit(net, qft_circuit(nq))
    return net, vecs


def cfg2_network(nq=24, depth=20, seed=None):
    """Config 2: 24-qubit brickwork, depth 20, ket-0 input, seeded output bitstring."""
    rng = np.random.default_rng(cfg_seed(2) if seed is None else seed)
    gates = brickwork_gates(nq, depth, rng)
    bits = rng.integers(0, 2, size=nq)
    return amplitude_network(nq, gates, bits), gates, bits


def cfg3_network(rows=6, cols=6, cycles=16, seed=None):
    """Config 3: rows x cols grid, ABCD coupler pattern, ket-0 input, seeded bitstring."""
    rng = np.random.default_rng(cfg_seed(3) if seed is None else seed)
    gates = rqc2d_gates(rows, cols, cycles, rng)
    bits = rng.integers(0, 2, size=rows * cols)
    return amplitude_network(rows * cols, gates, bits), gates, bits


def notebook_expectation_network(N=20, seed=None, is_decompose=False, cgc=None):
    """``expectation_value(cgc)`` of examples/expectation_value_optimization_example.ipynb (cells 2-12): the closed
    network <random bond-2 MPS| circuit |same MPS> behind the reference's only published timings.  Restated from the
    notebook's helper cell: ``ClosedMPS`` with *reversed* open legs, ``crand`` uniform in the unit square, the bra =
    the ket's tensors pushed in reverse site order, bond contractions = the ket's shifted by the tensor count before
    the push.  Same seeded stream as the product's generator (circuits.py of the package)."""
    rng = np.random.default_rng(cfg_seed(6) if seed is None else seed)

    def crand(*dims):
        return rng.random(dims) + 1j * rng.random(dims)
    t0 = [Tensor(crand(2, 2))] + [Tensor(crand(2, 2, 2)) for _ in range(2, N)] + [Tensor(crand(2, 2))]
    cons0 = [Summation([(1, 2), (2, 1)])] + [Summation([(i, 3), (i + 1, 1)]) for i in range(2, N)]
    open0 = list(reversed([(1, 1)] + [(i, 2) for i in range(2, N + 1)]))
    net = Network(list(t0), list(cons0), list(open0))
    tensor_circuit(net, qft_circuit(N) if cgc is None else cgc, is_decompose=is_decompose)
    step = len(net.tensors)
    net.contractions = net.contractions + [Summation([(t + step, l) for (t, l) in s.idx]) for s in cons0]
    for i in range(1, N + 1):
        net.tensors.append(t0[N - i])
        net.contractions.append(Summation([net.openidx[-1], (len(net.tensors), open0[N - i][1])]))
        net.openidx.pop()
    return net
