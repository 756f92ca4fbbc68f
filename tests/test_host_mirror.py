"""Host-side mirror of the reference surface (no GPU): constructors, error strings, symbolic
network growth and synthetic generators must agree with the oracle restatement."""
import numpy as np
import pytest

from conftest import to_oracle
from oracle import circuits as ocirc
from oracle import gates as og
from oracle import mpo as ompo
from oracle import network as on


def nets_equal(net, onet):
    assert len(net.tensors) == len(onet.tensors)
    for a, b in zip(net.tensors, onet.tensors):
        assert a.data.shape == b.data.shape and np.array_equal(a.data, b.data)
    assert [s.idx for s in net.contractions] == [s.idx for s in onet.contractions]
    assert list(net.openidx) == list(onet.openidx)


def test_helpers_and_tensor(q):  # test/test_helper.jl, test/test_tensor_circuit.jl:7-19
    assert q.shift_pair((1, 1), 5) == (6, 1)
    assert q.shift_summation(q.Summation([(1, 1), (6, 1)]), 5) == q.Summation([(6, 1), (11, 1)])
    assert q.is_power_two(1024) and not q.is_power_two(1023) and not q.is_power_two(0)
    d = np.arange(12.0).reshape(2, 6) + 0j
    t = q.Tensor(d)
    assert t.reshape(3, 4).isapprox(t.reshape((3, 4))) and t.reshape(3, 4).size() == (3, 4)
    assert np.array_equal(t.reshape(3, 4).data, np.reshape(d, (3, 4), order="F"))
    assert np.array_equal(t.transpose().data, d.T)


def test_gate_set_matches_oracle(q):
    g = q.gates
    for a, b in ((g.X, og.X), (g.Y, og.Y), (g.Z, og.Z), (g.HadamardGate, og.H), (g.SGate, og.S), (g.TGate, og.T),
                 (g.SwapGate, og.SWAP), (g.PhaseShiftGate(0.3), og.phase_shift(0.3)), (g.RxGate(0.4), og.rx(0.4)),
                 (g.RyGate(0.5), og.ry(0.5)), (g.RzGate(0.6), og.rz(0.6)), (g.ControlledGate(g.X), og.controlled(og.X))):
        assert np.array_equal(a, b)
    Ucnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)  # test/test_mpo.jl:81
    assert np.array_equal(q.circuit_gate(1, g.X, 2).matrix, Ucnot)
    for a, b in zip(q.qft_circuit(5), og.qft_circuit(5)):
        assert a.iwire == b.iwire and np.array_equal(a.matrix, b.matrix)
    with pytest.raises(ValueError, match="Repeated wires"):
        q.CircuitGate((1, 1), np.eye(4))


def test_generators_match_oracle(q):
    for (pn, on_) in ((q.circuits.cfg2_network(8, 5, seed=3), ocirc.cfg2_network(8, 5, seed=3)),
                      (q.circuits.cfg3_network(3, 3, 6, seed=4), ocirc.cfg3_network(3, 3, 6, seed=4))):
        nets_equal(pn[0], on_[0])
        assert list(pn[2]) == list(on_[2])
    pn, pv = q.circuits.cfg1_qft_network(5)
    onet, ov = ocirc.cfg1_qft_network(5)
    nets_equal(pn, onet)
    assert np.array_equal(pv, ov)
    net = q.circuits.cfg2_network()[0]
    assert len(net.tensors) == 278 and len(net.contractions) == 484 and not net.openidx
    net = q.circuits.cfg3_network()[0]
    assert len(net.tensors) == 312 and len(net.contractions) == 516
    rng = np.random.default_rng(0)
    U = q.circuits.haar_unitary(4, rng)
    assert np.abs(U.conj().T @ U - np.eye(4)).max() < 1e-14


def test_tensor_circuit_structure_matches_oracle(q):  # src/tensor_circuit.jl:44-51 (non-decomposed)
    rng = np.random.default_rng(1)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    ts = [r(2, 6), r(2, 6, 7), r(2, 7)]
    net = q.GeneralTensorNetwork([q.Tensor(t) for t in ts], [q.Summation([(1, 2), (2, 2)]), q.Summation([(2, 3), (3, 2)])],
                                 [(1, 1), (2, 1), (3, 1)])
    onet = to_oracle(net)
    q.tensor_circuit(net, q.qft_circuit(3))
    ompo.tensor_circuit(onet, og.qft_circuit(3))
    nets_equal(net, onet)
    assert q.contract_rep(net) == __import__("oracle.contract", fromlist=["x"]).contract_rep(onet)
    with pytest.raises(AssertionError):
        q.tensor_circuit(net, q.circuit_gate(7, q.gates.X))


def test_mps_constructors_and_errors(q):  # test/test_mps.jl:21-32, 90-145
    rng = np.random.default_rng(2)
    T3 = q.Tensor(rng.standard_normal((2, 2, 2)))
    T2 = q.Tensor(rng.standard_normal((2, 2)))
    o = q.OpenMPS(T3, 4)
    assert o.openidx == [(1, 1), (1, 2), (2, 2), (3, 2), (4, 2), (4, 3)]
    assert [s.idx for s in o.contractions] == [[(1, 3), (2, 1)], [(2, 3), (3, 1)], [(3, 3), (4, 1)]]
    c = q.ClosedMPS(T2, T3, T2, 4)
    assert c.openidx == [(1, 1), (2, 2), (3, 2), (4, 2)] and c.contractions[0] == q.Summation([(1, 2), (2, 1)])
    p = q.PeriodicMPS(T3, 3)
    assert p.contractions[-1] == q.Summation([(3, 3), (1, 1)]) and p.openidx == [(1, 2), (2, 2), (3, 2)]
    with pytest.raises(ValueError, match="Tensors must have 3 legs"):
        q.OpenMPS([T3, T2, T3])
    with pytest.raises(ValueError, match="First tensor must have 2 legs"):
        q.ClosedMPS([T3, T3, T2])
    with pytest.raises(ValueError, match="except the first and last one"):
        q.ClosedMPS([T2, T2, T2])
    with pytest.raises(ValueError, match="Last tensor must have 2 legs"):
        q.ClosedMPS([T2, T3, T3])
    with pytest.raises(ValueError, match="first leg must contract with last leg"):
        q.MPS([T3, T3], [q.Summation([(1, 2), (2, 1)])], [(1, 1)])
    with pytest.raises(ValueError, match="last leg must contract with first leg"):
        q.MPS([T3, T3], [q.Summation([(1, 3), (2, 2)])], [(1, 1)])
    with pytest.raises(ValueError, match="can only have 2 or 3 legs"):
        q.MPS([T3, q.Tensor(np.zeros((2, 2, 2, 2)))], [q.Summation([(1, 3), (2, 1)])], [(1, 1)])
    cp = o.copy()
    cp.tensors[0] = T2
    assert o.tensors[0] is T3 and cp.contractions == o.contractions       # shallow copy (src/mps.jl:203)
    with pytest.raises(ValueError, match="Permutation order cannot contain repeat values"):
        q.permute(o, [1, 1, 2, 3])
    with pytest.raises(ValueError, match="same length"):
        q.permute(o, [1, 2])
    with pytest.raises(ValueError, match="can only contain positive"):
        q.permute(o, [0, 1, 2, 3])
    with pytest.raises(ValueError, match="cannot exceed number of wires"):
        q.permute(o, [1, 2, 3, 7])
    with pytest.raises(ValueError, match="must be positive"):
        q.switch(o, 0, 1)
    with pytest.raises(ValueError, match="less than or equal to the number of open wires"):
        q.switch(o, 1, 9)


def test_mpo_argument_errors(q):  # test/test_mpo.jl:124-140 (no SVD is reached)
    with pytest.raises(ValueError, match="Direct conversion to MPS form is not support"):
        q.MPO([], [], [])
    with pytest.raises(ValueError, match="Need at least one qubit"):
        q.MPO(np.ones((1, 1)))
    one = q.MPO(np.array([[0, 1], [1, 0]], dtype=complex))       # one-qubit operator: no SVD needed
    assert one.openidx == [(1, 2), (1, 1)] and len(one.tensors) == 1
    psi = q.circuits.amplitude_network(3, [], None)
    with pytest.raises(ValueError, match="Repeated wires are not valid"):
        q.apply_MPO(psi, np.eye(4), (1, 1))
    with pytest.raises(ValueError, match="Wires must be positive integers"):
        q.apply_MPO(psi, np.eye(4), (0, 1))
    with pytest.raises(ValueError, match="between 1 and n"):
        q.apply_MPO(psi, one, (5,))
    out = q.apply_MPO(psi, one, (2,))
    oout = ompo.apply_MPO(to_oracle(psi), ompo.MPO(np.array([[0, 1], [1, 0]], dtype=complex)), (2,))
    nets_equal(out, oout)
    with pytest.raises(ValueError, match="Wires not sorted"):
        q.extend_MPO(one, (1, 3))


def test_packed_pointer_table_round_trip(q):
    """_lib.packed_ptrs (the marshalling of `contract(net)` for many small tensors): every table entry must point at
    the column-major image of its tensor."""
    import ctypes as C
    from qaintensor_b200 import _lib
    rng = np.random.default_rng(3)
    shapes = [(2,), (2, 2), (2, 2, 2), (2, 2, 2, 2), (3, 1, 2), (1,), (4, 4)] * 6
    for dtype in (np.complex128, np.complex64):
        arrs = [np.asfortranarray((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(dtype)) for s in shapes]
        keep, ptrs = _lib.packed_ptrs(arrs)
        for i, a in enumerate(arrs):
            raw = (C.c_char * a.nbytes).from_address(ptrs[i])
            back = np.frombuffer(raw, dtype=dtype).reshape(a.shape, order="F")
            assert np.array_equal(back, a)
