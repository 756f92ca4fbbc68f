"""The library's own network builder (csrc/network.cpp, ``qtn_net_*``) against the Python mirror of the
reference's symbolic code (structure bit-exact) and, on the GPU, against the CPU oracle (amplitudes 1e-10)."""
import numpy as np
import pytest

from conftest import random_TN, rel_err, to_oracle
from oracle import contract as oc


def _same(a, b):
    return (a.contractions == b.contractions and list(a.openidx) == list(b.openidx) and len(a.tensors) == len(b.tensors) and
            all(x.data.shape == y.data.shape and np.array_equal(x.data, y.data) for x, y in zip(a.tensors, b.tensors)))


def _ket_network(q, nq):
    return q.GeneralTensorNetwork([q.Tensor(np.array([1, 0], dtype=np.complex128)) for _ in range(nq)], [],
                                  [(i, 1) for i in range(1, nq + 1)])


def test_round_trip_and_tensor_circuit_structure(q):
    rng = np.random.default_rng(2)
    net = random_TN(q, 10, 20, rng)
    assert _same(q.NativeNetwork.from_network(net).to_network(), net)
    for nq, gates in ((12, q.qft_circuit(12)), (24, q.circuits.brickwork_gates(24, 20, rng)),
                      (36, q.circuits.rqc2d_gates(6, 6, 16, rng))):
        want = _ket_network(q, nq)
        q.tensor_circuit(want, gates)
        nn = q.NativeNetwork.from_network(_ket_network(q, nq))
        nn.tensor_circuit(gates)
        assert nn.sizes() == (len(want.tensors), len(want.contractions), nq)
        assert _same(nn.to_network(), want)
    # amplitude network of BASELINE config 3, and both orderings, bit-exact with the mirror
    want, gates, bits = q.circuits.cfg3_network()
    nn = q.NativeNetwork.from_network(_ket_network(q, 36))
    nn.tensor_circuit(gates)
    nn.close_wires(bits)
    assert _same(nn.to_network(), want)
    nn.optimize_contraction_order()
    q.optimize_contraction_order(want)
    assert _same(nn.to_network(), want)
    nn.optimize_contraction_order(method="search", ntrials=16, seed=5, max_log2_elems=31)
    q.optimize_contraction_order(want, method="search", ntrials=16, seed=5, max_log2_elems=31)
    assert _same(nn.to_network(), want)


def test_apply_mpo_structure_and_error_strings(q):
    from qaintensor_b200 import mpo as pm
    rng = np.random.default_rng(3)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    # a hand-built 3-site operator network in the MPO layout of src/mpo.jl:57-88 (no SVD needed on the host)
    op = pm.MPO.__new__(pm.MPO)
    op.tensors = [q.Tensor(r(2, 2, 3)), q.Tensor(r(3, 2, 2, 4)), q.Tensor(r(4, 2, 2))]
    op.contractions = [q.Summation([(1, 3), (2, 1)]), q.Summation([(2, 4), (3, 1)])]
    op.openidx = [(3, 2), (2, 2), (1, 1), (3, 3), (2, 3), (1, 2)]
    psi = _ket_network(q, 5)
    q.tensor_circuit(psi, q.circuits.brickwork_gates(5, 3, rng))
    for iwire in ((1, 2, 3), (2, 4, 5), (5, 3, 1)):
        want = q.apply_MPO(psi, op, iwire)
        got = q.NativeNetwork.from_network(psi).apply_mpo(q.NativeNetwork.from_network(op), iwire)
        assert _same(got.to_network(), want)
    npsi, nop = q.NativeNetwork.from_network(psi), q.NativeNetwork.from_network(op)
    with pytest.raises(q.QtnError, match="Repeated wires are not valid."):
        npsi.apply_mpo(nop, (1, 1, 2))
    with pytest.raises(q.QtnError, match="Wires must be integers between 1 and n"):
        npsi.apply_mpo(nop, (1, 2, 6))
    with pytest.raises(q.QtnError, match="2 \\* 2 open legs"):
        npsi.apply_mpo(nop, (1, 2))
    g = q.CircuitGate((1, 7), np.eye(4))
    with pytest.raises(q.QtnError, match="more wires than the network has open legs"):
        npsi.tensor_circuit([g])
    with pytest.raises(q.QtnError, match="Repeated wires"):
        npsi.tensor_circuit([_FakeGate((2, 2), np.eye(4))])  # CircuitGate's own constructor already rejects this
    bad = q.GeneralTensorNetwork([q.Tensor(r(2, 2))], [], [(1, 1)])
    with pytest.raises(q.QtnError, match="tensor leg without contraction or open index"):
        q.NativeNetwork.from_network(bad).optimize_contraction_order(method="search")


def test_extend_mpo_structure_and_oracle_cross_check(q):
    from oracle import mpo as ompo, network as on
    from qaintensor_b200 import mpo as pm
    rng = np.random.default_rng(7)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731

    def hand_mpo(bonds):
        """Operator network in the MPO layout of src/mpo.jl:57-88 with the given inner bond dimensions."""
        M = len(bonds) + 1
        op = pm.MPO.__new__(pm.MPO)
        op.tensors = [q.Tensor(r(2, 2, bonds[0]))] + [q.Tensor(r(bonds[i - 1], 2, 2, bonds[i])) for i in range(1, M - 1)] + \
                     [q.Tensor(r(bonds[-1], 2, 2))]
        op.contractions = [q.Summation([(i, 3 if i == 1 else 4), (i + 1, 1)]) for i in range(1, M)]
        op.openidx = [(M - i + 1, 2) for i in range(1, M)] + [(1, 1)] + [(M - i + 1, 3) for i in range(1, M)] + [(1, 2)]
        return op

    nfail = 0
    for bonds, iwire in (([3], (4, 1)), ([2, 4], (6, 4, 1)), ([4, 3], (5, 4, 2)), ([2], (3, 1)), ([2, 2], (4, 3, 1)), ([3, 2, 2], (5, 4, 3, 1))):
        want = hand_mpo(bonds)
        nn = q.NativeNetwork.from_network(want)
        ocopy = object.__new__(ompo.MPO)
        on.Network.__init__(ocopy, [on.Tensor(t.data) for t in want.tensors], [on.Summation(c.idx) for c in want.contractions],
                            list(want.openidx))
        try:
            q.extend_MPO(want, iwire)      # Python mirror (mutates)
        except IndexError:                 # the reference indexes past the tensor list for this wire pattern (BoundsError there)
            with pytest.raises(IndexError):
                ompo.extend_MPO(ocopy, iwire)
            with pytest.raises(q.QtnError, match="pipe position out of range"):
                nn.extend_mpo(iwire)
            nfail += 1
            continue
        nn.extend_mpo(iwire)               # C++ builder (mutates)
        ompo.extend_MPO(ocopy, iwire)      # oracle (mutates)
        got = nn.to_network()
        assert _same(got, want)
        assert [c.idx for c in got.contractions] == [c.idx for c in ocopy.contractions] and list(got.openidx) == list(ocopy.openidx)
        assert all(np.array_equal(a.data, b.data) for a, b in zip(got.tensors, ocopy.tensors))
    assert 0 < nfail < 6
    nn = q.NativeNetwork.from_network(hand_mpo([3]))
    for iwire, msg in (((2, 1), "MPO is already decomposed in N tensors"), ((1, 4), "Wires not sorted"), ((3, 3), "Repeated wires are not valid."),
                       ((2, 0), "Wires must be positive integers."), ((5, 3, 1), "MPO length does not match the wires")):
        with pytest.raises(q.QtnError, match=msg):
            nn.extend_mpo(iwire)


class _FakeGate:
    def __init__(self, iwire, matrix):
        self.iwire, self.matrix = iwire, matrix


@pytest.mark.gpu
def test_native_network_contract_gpu(gpu):
    q = gpu
    rng = np.random.default_rng(4)
    # open legs: full state vector of a 10-qubit brickwork circuit
    gates = q.circuits.brickwork_gates(10, 6, rng)
    want_net = q.circuits.amplitude_network(10, gates, None)
    nn = q.NativeNetwork.from_network(_ket_network(q, 10))
    nn.tensor_circuit(gates)
    want = oc.contract(to_oracle(want_net))
    got = nn.contract()
    assert got.shape == want.shape and rel_err(got, want) < 1e-10
    assert rel_err(nn.contract(precision="c64"), want) < 1e-4
    # closed amplitude, reference order / searched order / sliced
    net, gates, bits = q.circuits.cfg2_network(16, 12, seed=11)
    want = complex(oc.contract(to_oracle(net)))
    nn = q.NativeNetwork.from_network(_ket_network(q, 16))
    nn.tensor_circuit(gates)
    nn.close_wires(bits)
    nn.optimize_contraction_order()
    assert abs(complex(nn.contract()) - want) < 1e-10 * abs(want)
    assert abs(complex(nn.contract(max_log2_elems=8)) - want) < 1e-10 * abs(want)
    nn.optimize_contraction_order(method="search", ntrials=32)
    assert abs(complex(nn.contract()) - want) < 1e-10 * abs(want)
    # single tensor: permutedims by the open legs (src/contract.jl:243-245)
    d = rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))
    one = q.NativeNetwork.from_network(q.GeneralTensorNetwork([q.Tensor(d)], [], [(1, 2), (1, 1)]))
    assert np.array_equal(one.contract(), d.T)
