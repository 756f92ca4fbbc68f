"""Synthetic workloads of BASELINE.json (EXTENSION: the reference has no generators).

Seeds follow SURVEY.md section 8(d): ``20261017 + 1000*cfg + index``; Haar 2-qubit gates
are QR of a complex Ginibre matrix with the phases of R's diagonal fixed.
"""
import numpy as np

from .gates import CircuitGate, qft_circuit
from .tensor_circuit import tensor_circuit
from .tensor_network import GeneralTensorNetwork, Summation, Tensor


def cfg_seed(cfg, index=0):
    return 20261017 + 1000 * cfg + index


def haar_unitary(n, rng):
    z = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(2.0)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def brickwork_gates(nq, depth, rng):
    return [CircuitGate((q, q + 1), haar_unitary(4, rng))
            for layer in range(depth) for q in range(1 + layer % 2, nq, 2)]


def rqc2d_gates(rows, cols, cycles, rng):
    """Coupler classes A (horizontal, even columns), B (vertical, even rows), C (horizontal,
    odd columns), D (vertical, odd rows), repeating; qubit (r, c) is wire r*cols + c + 1."""
    gates = []
    for cyc in range(cycles):
        kind = cyc % 4
        if kind % 2 == 0:
            pairs = [(r * cols + c + 1, r * cols + c + 2) for r in range(rows) for c in range(kind // 2, cols - 1, 2)]
        else:
            pairs = [(r * cols + c + 1, (r + 1) * cols + c + 1) for r in range(kind // 2, rows - 1, 2) for c in range(cols)]
        gates += [CircuitGate(p, haar_unitary(4, rng)) for p in pairs]
    return gates


def amplitude_network(nq, gates, bits, input_vectors=None):
    """<bits| gates |input>: rank-1 kets, ``tensor_circuit!`` gates, rank-1 bras; no open legs."""
    kets = [np.array([1, 0], dtype=np.complex128) if input_vectors is None else np.asarray(input_vectors[i], np.complex128)
            for i in range(nq)]
    net = GeneralTensorNetwork([Tensor(k) for k in kets], [], [(i, 1) for i in range(1, nq + 1)])
    tensor_circuit(net, gates)
    if bits is not None:
        for w, b in enumerate(bits, 1):
            v = np.zeros(2, dtype=np.complex128)
            v[int(b)] = 1.0
            net.tensors.append(Tensor(v))
            net.contractions.append(Summation([net.openidx[w - 1], (len(net.tensors), 1)]))
        net.openidx = []
    return net


def cfg1_qft_network(nq=12, seed=None):
    rng = np.random.default_rng(cfg_seed(1) if seed is None else seed)
    vecs = rng.standard_normal((nq, 2)) + 1j * rng.standard_normal((nq, 2))
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    return amplitude_network(nq, qft_circuit(nq), None, vecs), vecs


def cfg2_network(nq=24, depth=20, seed=None):
    rng = np.random.default_rng(cfg_seed(2) if seed is None else seed)
    gates = brickwork_gates(nq, depth, rng)
    bits = rng.integers(0, 2, size=nq)
    return amplitude_network(nq, gates, bits), gates, bits


def cfg3_network(rows=6, cols=6, cycles=16, seed=None):
    rng = np.random.default_rng(cfg_seed(3) if seed is None else seed)
    gates = rqc2d_gates(rows, cols, cycles, rng)
    bits = rng.integers(0, 2, size=rows * cols)
    return amplitude_network(rows * cols, gates, bits), gates, bits


def notebook_expectation_network(N=20, seed=None, is_decompose=False, cgc=None):
    """The network the reference's only published timings are taken on
    (``examples/expectation_value_optimization_example.ipynb``, cells 2-12):
    <random bond-2 MPS| circuit |same MPS> with no open legs, circuit = ``qft_circuit(N)`` unless given.

    Follows the notebook's own helper cell, not src/mps.jl: its ``ClosedMPS`` lists the open legs in reverse
    (``reverse([1 => 1; [i => 2 for i in 2:l]])``), ``crand`` is uniform on [0, 1) + i [0, 1), the bra re-uses the
    ket's tensors (no conjugation) appended in reverse site order, and the bra's bond contractions are the ket's
    shifted by the tensor count *before* the bra tensors are pushed -- so they pair legs of the reversed list
    (all extents are 2, so the network is valid; it is reproduced as published, not corrected)."""
    rng = np.random.default_rng(cfg_seed(6) if seed is None else seed)

    def crand(*dims):
        return np.asfortranarray(rng.random(dims) + 1j * rng.random(dims))
    t0 = [Tensor(crand(2, 2))] + [Tensor(crand(2, 2, 2)) for _ in range(2, N)] + [Tensor(crand(2, 2))]
    cons0 = [Summation([(1, 2), (2, 1)])] + [Summation([(i, 3), (i + 1, 1)]) for i in range(2, N)]
    open0 = list(reversed([(1, 1)] + [(i, 2) for i in range(2, N + 1)]))
    net = GeneralTensorNetwork(list(t0), list(cons0), list(open0))
    tensor_circuit(net, qft_circuit(N) if cgc is None else cgc, is_decompose=is_decompose)
    step = len(net.tensors)
    net.contractions = net.contractions + [Summation([(t + step, l) for (t, l) in s.idx]) for s in cons0]
    for i in range(1, N + 1):
        net.tensors.append(t0[N - i])
        net.contractions.append(Summation([net.openidx[-1], (len(net.tensors), open0[N - i][1])]))
        net.openidx.pop()
    return net
