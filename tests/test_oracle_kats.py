"""Pins the CPU oracle against every literal known-answer the reference's tests hold for
this path (test/test_treewidth.jl, test/test_helper.jl, test/test_mpo.jl:81), against the
reference tests' cross-path relations (seeded), and against independent dense maths."""
import itertools
import warnings

import numpy as np
import pytest

from oracle import circuits as ocirc
from oracle import contract as oc
from oracle import gates as og
from oracle import mpo as ompo
from oracle import mps as omps
from oracle import network as on
from oracle import network2graph as o2g
from oracle import plan as oplan
from oracle import svd as osvd
from oracle.lightgraphs import Graph, complete_graph


def four_cycle(rng):  # test/test_treewidth.jl:41-51
    A = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    return on.Network([on.Tensor(A.copy()) for _ in range(4)],
                      [on.Summation([(1, 2), (2, 1)]), on.Summation([(2, 2), (3, 1)]),
                       on.Summation([(3, 2), (4, 1)]), on.Summation([(4, 2), (1, 1)])], []), A


def test_helper_shifts():  # test/test_helper.jl:6-25
    assert on.shift_pair((1, 1), 5) == (6, 1)
    assert on.shift_summation(on.Summation([(1, 1), (6, 1)]), 5) == on.Summation([(6, 1), (11, 1)])
    for N in range(2, 11):
        assert on.is_power_two(2 ** N) and not on.is_power_two(2 ** N - 1)


def test_subset():  # test/test_treewidth.jl:53-69
    A, B = [1, 2, 3], [1, 2, 3, 4]
    assert o2g.subset(A, B) and not o2g.subset(B, A)
    assert o2g.subset((1, 2), A) and o2g.subset(A, (1, 2, 3, 4)) and o2g.subset(1, A) and o2g.subset(range(1, 4), A)
    assert o2g.subset("ba", "abc")


def test_local_circuit_graph():  # test/test_treewidth.jl:83-98
    G = Graph(4)
    for e in [(1, 2), (2, 3), (3, 4)]:
        G.add_edge(*e)
    assert G == o2g.local_circuit_graph(4, 2)
    G.add_edge(1, 3); G.add_edge(2, 4)
    assert G == o2g.local_circuit_graph(4, 3)
    G.add_edge(1, 4)
    assert G == o2g.local_circuit_graph(4, 4)


def test_random_graph_counts():  # test/test_treewidth.jl:71-81
    G = o2g.random_graph(10, 20, np.random.default_rng(0))
    assert G.nv() == 10 and G.ne() == 20
    with pytest.raises(ValueError, match="Number of edges must be smaller or equal"):
        o2g.random_graph(10, 46, np.random.default_rng(0))


def test_network_and_line_graph_four_cycle():  # test/test_treewidth.jl:100-125 + SURVEY App. B
    net, _ = four_cycle(np.random.default_rng(1))
    G, _ = o2g.network_graph(net)
    assert G.nv() == 4 and all(d == 2 for d in G.degree())
    LG, _ = o2g.line_graph_of_graph(G)
    assert LG.nv() == 4 and all(d == 2 for d in LG.degree())
    LG0, nodeinfo = o2g.line_graph(net)
    assert LG0.nv() == 4 and all(d == 2 for d in LG0.degree())
    assert nodeinfo == [(1, 2, 1), (1, 4, 4), (2, 3, 2), (3, 4, 3)]
    assert LG0.adj == [[2, 3], [1, 4], [1, 4], [2, 3]]
    bad = net.copy()
    bad.contractions = [on.Summation([(1, 2), (2, 1), (2, 2), (3, 1)]), on.Summation([(3, 2), (4, 1), (4, 2), (1, 1)])]
    with pytest.raises(ValueError, match="Contractions of more than 2 tensors not supported"):
        o2g.network_graph(bad)
    op = net.copy()
    c = op.contractions.pop()
    op.openidx += [c.idx[0], c.idx[1]]
    with pytest.warns(UserWarning, match="All open indices are disregarded"):
        o2g.line_graph(op)


def test_hand_traced_order_vector():  # SURVEY.md Appendix B (regression KAT)
    net, A = four_cycle(np.random.default_rng(2))
    LG, _ = o2g.line_graph(net)
    assert o2g.min_fill_ordering(LG) == [1, 4, 2, 3]
    tw, tree, bags = o2g.tree_decomposition(LG)
    assert tw == 2 and bags == [[2, 3, 4], [2, 3, 1]] and tree.adj == [[2], [1]]
    assert o2g.contraction_order(net) == [(3, 4, 3), (1, 4, 4), (2, 3, 2), (1, 2, 1)]
    n2 = net.copy()
    assert o2g.optimize_contraction_order(n2) == [3, 4, 2, 1]
    assert set(n2.contractions) == set(net.contractions) and n2.tensors == net.tensors
    v = np.trace(A @ A @ A @ A)
    assert abs(oc.contract(net) - v) < 1e-13 * abs(v) and abs(oc.contract(n2) - v) < 1e-13 * abs(v)


def hand_traced():
    import json
    import os
    k = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hand_traced_orders.json")))
    return {n: v for n, v in k.items() if not n.startswith("_")}


@pytest.mark.parametrize("name", sorted(hand_traced()))
def test_hand_traced_orders_all_stages(name):
    """tests/golden/hand_traced_orders.json: four networks (4-cycle, self-contraction, parallel edges, a 5-cycle whose
    line graph needs two fill edges) traced by hand through src/network2graph.jl -- every intermediate stage."""
    k = hand_traced()[name]
    rng = np.random.default_rng(3)
    net = on.Network([on.Tensor(rng.standard_normal((2,) * l) + 0j) for l in k["legs"]],
                     [on.Summation([tuple(p) for p in c]) for c in k["contractions"]], [])
    LG, nodeinfo = o2g.line_graph(net)
    assert [list(t) for t in nodeinfo] == k["nodeinfo"] and LG.adj == k["lg_adj"]
    assert o2g.min_fill_ordering(LG) == k["min_fill"]
    tw, _, bags = o2g.tree_decomposition(LG)
    assert tw == k["tw"] and bags == k["bags"]
    assert [list(t) for t in o2g.contraction_order(net)] == k["order"]
    n2 = net.copy()
    assert o2g.optimize_contraction_order(n2) == k["perm"]
    v = oc.contract(net)
    assert abs(oc.contract(n2) - v) <= 1e-12 * abs(v)      # test/test_treewidth.jl:327: the value is order-invariant


def test_nodeinfo_random_tn():  # test/test_treewidth.jl:129-139
    rng = np.random.default_rng(3)
    cons = []
    nlegs = [0] * 10
    for _ in range(20):
        n1 = int(rng.integers(1, 10)); n2 = int(rng.integers(n1 + 1, 11))
        nlegs[n1 - 1] += 1; nlegs[n2 - 1] += 1
        cons.append(on.Summation([(n1, nlegs[n1 - 1]), (n2, nlegs[n2 - 1])]))
    net = on.Network([on.Tensor(np.zeros((2,) * max(n, 1))) for n in nlegs], cons, [])
    _, nodeinfo = o2g.line_graph(net)
    expect = {tuple(sorted((c.idx[0][0], c.idx[1][0]))) + (k,) for k, c in enumerate(cons, 1)}
    assert set(nodeinfo) == expect


def test_interaction_graph_qft():  # test/test_treewidth.jl:149-156
    assert og.interaction_graph(og.qft_circuit(10)) == complete_graph(10)


def test_lacking_and_rem_vertex_fill():  # test/test_treewidth.jl:159-178
    G = complete_graph(5)
    assert o2g.lacking_for_clique_neigh(G, 1) == (0, [])
    G.rem_edge(2, 3)
    assert o2g.lacking_for_clique_neigh(G, 1) == (1, [(2, 3)])
    ordering, vl = [], [1, 2, 3, 4, 5]
    o2g.rem_vertex_fill(G, 1, [(2, 3)], ordering, vl)
    assert G == complete_graph(4) and ordering == [1] and vl == [5, 2, 3, 4]


def test_treewidth_known_answers():  # test/test_treewidth.jl:204-221
    for n in (10, 25, 50):
        assert o2g.tree_decomposition(complete_graph(n))[0] == n - 1
    for k in range(2, 6):
        assert o2g.tree_decomposition(o2g.local_circuit_graph(10, k))[0] == k - 1


def test_is_tree_decomposition_example():  # test/test_treewidth.jl:223-277
    G = Graph(5)
    for e in [(1, 2), (1, 4), (2, 3), (4, 3), (5, 3), (4, 5)]:
        G.add_edge(*e)
    tree = Graph(3)
    tree.add_edge(1, 2); tree.add_edge(1, 3)
    assert o2g.is_tree_decomposition(G, tree, [[2, 3, 4], [2, 4, 1], [3, 4, 5]])
    with pytest.warns(UserWarning, match="Union of bags is not equal to union of vertices"):
        assert not o2g.is_tree_decomposition(G, tree, [[2, 3, 4], [2, 4], [3, 4, 5]])
    with pytest.warns(UserWarning, match=r"Edge \(2, 3\) not found in any bag"):
        assert not o2g.is_tree_decomposition(G, tree, [[3, 4], [2, 4, 1], [3, 4, 5]])
    with pytest.warns(UserWarning, match="Subgraph for vertex 4 not connected"):
        assert not o2g.is_tree_decomposition(G, tree, [[2, 3], [2, 4, 1], [3, 4, 5]])


def test_tree_decomposition_random_graphs_valid():  # test/test_treewidth.jl:280-297
    rng = np.random.default_rng(4)
    for _ in range(3):
        Nn = int(rng.integers(20, 51))
        Ne = int(rng.integers(3 * Nn, Nn * (Nn - 1) // 2 + 1))
        G = o2g.random_graph(Nn, Ne, rng)
        tw, tree, bags = o2g.tree_decomposition(G)
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            assert o2g.is_tree_decomposition(G, tree, bags)


def test_contraction_order_is_permutation_and_errors():  # test/test_treewidth.jl:300-316
    net, _ = four_cycle(np.random.default_rng(5))
    H, edges = o2g.line_graph(net)
    order = o2g.contraction_order_graph(H, edges)
    assert sorted(e[2] for e in order) == list(range(1, len(edges) + 1))
    edges.pop()
    with pytest.raises(ValueError, match="Invalid list of edges for `H`"):
        o2g.contraction_order_graph(H, edges)


def dft_bitrev(N):
    n = 1 << N
    F = np.exp(2j * np.pi * np.outer(np.arange(n), np.arange(n)) / n) / np.sqrt(n)
    rev = np.array([int(format(i, "0%db" % N)[::-1], 2) for i in range(n)])
    return F[rev][:, rev]


def test_qft_circuit_is_the_dft():  # independent maths for the un-vendored Qaintmodels.qft_circuit
    for N in (3, 5, 6):
        n = 1 << N
        U = np.stack([og.apply(np.eye(n)[:, i], og.qft_circuit(N)) for i in range(n)], axis=1)
        assert np.abs(U - dft_bitrev(N)).max() < 1e-13


def test_cfg1_qft12_state_vector_vs_dft():  # BASELINE config 1 (CPU-runnable case)
    net, vecs = ocirc.cfg1_qft_network(12)
    psi0 = vecs[0]
    for v in vecs[1:]:
        psi0 = np.kron(v, psi0)
    stats = []
    out = oc.contract(net, stats=stats).reshape(-1, order="F")
    ref = dft_bitrev(12) @ psi0
    assert np.abs(out - ref).max() < 1e-12
    assert len(net.tensors) == 96 and len(net.contractions) == 156 and len(stats) == 95
    assert 8e6 < sum(8 * m * n * k for m, n, k in stats) < 1e7  # ~8.4e6 flop


def mps_like_network(rng):  # test/test_tensor_circuit.jl:36-47
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    return on.Network([on.Tensor(r(2, 6)), on.Tensor(r(2, 6, 7)), on.Tensor(r(2, 7))],
                      [on.Summation([(1, 2), (2, 2)]), on.Summation([(2, 3), (3, 2)])], [(1, 1), (2, 1), (3, 1)])


def test_tensor_circuit_qft3_and_exhaustive_order():  # test/test_tensor_circuit.jl:32-60
    psi = mps_like_network(np.random.default_rng(6))
    cgc = og.qft_circuit(3)
    ref = og.apply(oc.contract(psi).reshape(-1, order="F"), cgc)
    ompo.tensor_circuit(psi, cgc)
    assert np.allclose(oc.contract(psi).reshape(-1, order="F"), ref, rtol=1e-12, atol=1e-13)
    assert np.allclose(oc.contract(psi, True).reshape(-1, order="F"), ref, rtol=1e-12, atol=1e-13)


def test_single_tensor_contract():  # test/test_tensor_circuit.jl:21-31
    d = np.random.default_rng(7).standard_normal((2, 6)) + 0j
    net = on.Network([on.Tensor(d)], [], [(1, 1), (1, 2)])
    assert np.array_equal(oc.contract(net), d)
    net = on.Network([on.Tensor(d)], [], [(1, 2), (1, 1)])
    assert np.array_equal(oc.contract(net), d.T)


def test_decomposed_gates_vs_apply():  # test/test_tensor_circuit.jl:62-126
    rng = np.random.default_rng(8)
    psi = mps_like_network(rng)
    cgc = [og.circuit_gate(3, og.X, 1), og.circuit_gate(3, og.Y, 1), og.circuit_gate(1, og.Y, 2), og.circuit_gate(2, og.Z, 1)]
    ref = og.apply(oc.contract(psi).reshape(-1, order="F"), cgc)
    ompo.tensor_circuit(psi, cgc, is_decompose=True)
    assert np.allclose(oc.contract(psi).reshape(-1, order="F"), ref, rtol=1e-12, atol=1e-12)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    psi = on.Network([on.Tensor(r(2, 6)), on.Tensor(r(2, 6, 7)), on.Tensor(r(2, 7, 5)), on.Tensor(r(2, 5))],
                     [on.Summation([(1, 2), (2, 2)]), on.Summation([(2, 3), (3, 2)]), on.Summation([(3, 3), (4, 2)])],
                     [(1, 1), (2, 1), (3, 1), (4, 1)])
    cgc = [og.circuit_gate((4, 1), og.SWAP, 3), og.circuit_gate(2, og.Y), og.circuit_gate(3, og.X, 1),
           og.circuit_gate(2, og.Y), og.circuit_gate(4, og.X, (1, 2)), og.circuit_gate(3, og.Z)]
    ref = og.apply(oc.contract(psi).reshape(-1, order="F"), cgc)
    ompo.tensor_circuit(psi, cgc, is_decompose=True)
    assert np.allclose(oc.contract(psi).reshape(-1, order="F"), ref, rtol=1e-12, atol=1e-12)


def test_contract_svd_relations_and_errors():  # test/test_svd.jl:7-50
    rng = np.random.default_rng(9)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    t1, t2 = r(4, 4), r(4, 4)
    assert np.allclose(osvd.contract_svd(t1, t2, (2, 1)), t1 @ t2, rtol=1e-12, atol=1e-12)
    t1, t2 = r(2, 3, 4, 6), r(1, 5, 2, 4)
    tn1 = on.Network([on.Tensor(t1), on.Tensor(t2)], [on.Summation([(1, 1), (2, 3)])],
                     [(1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 4)])
    tn2 = on.Network([on.Tensor(t1), on.Tensor(t2)], [on.Summation([(1, 3), (2, 4)])],
                     [(1, 1), (1, 2), (1, 4), (2, 1), (2, 2), (2, 3)])
    assert np.allclose(oc.contract(tn1), osvd.contract_svd(t1, t2, (1, 3)), rtol=1e-12, atol=1e-12)
    assert np.allclose(oc.contract(tn2), osvd.contract_svd(t1, t2, (3, 4)), rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError, match="Error must be positive"):
        osvd.contract_svd(t1, t2, (1, 3), er=-0.3)
    with pytest.raises(ValueError, match="Dimensions of contraction legs do not match"):
        osvd.contract_svd(t1, r(8, 8), (1, 3))


def test_truncation_rule_and_bound():  # test/test_svd.jl:53-83 + SURVEY App. A.3
    s = np.exp(-np.arange(100.0))
    rng = np.random.default_rng(10)
    qs = [np.linalg.qr(rng.standard_normal((100, 100)))[0] for _ in range(4)]
    T1 = qs[0] @ np.diag(s) @ qs[1]
    T2 = qs[2] @ np.diag(s) @ qs[3]
    er = 1e-10
    mps = omps.ClosedMPS([on.Tensor(T1), on.Tensor(T2)])
    approx = omps.contract_svd_mps(mps, er=er)
    exact = oc.contract(mps)
    k = osvd.truncation_rank(s, er)
    tail = np.sqrt(np.cumsum(s[::-1] ** 2))
    assert k == 100 - (int(np.nonzero(tail > er)[0][0]) + 1) + 1 and 0 < k < 100
    n, nt = np.linalg.norm(s), np.linalg.norm(s[k:])
    assert np.linalg.norm(approx - exact) < 2 * n * nt + nt ** 2
    # strict '>' and exact zeros: er = 0 drops only exact zeros
    assert osvd.truncation_rank([3.0, 2.0, 0.0, 0.0], 0.0) == 2
    assert osvd.truncation_rank([3.0, 2.0, 1.0], 1.0) == 2          # tail == er is dropped
    assert osvd.truncation_rank([3.0, 2.0, 1.0], 0.999) == 3
    assert osvd.truncation_rank([0.0, 0.0], 0.0) == 0                 # reference throws here
    assert osvd.truncation_rank([3.0, 2.0, 1.0], 0.0, maxdim=2) == 2  # chi extension


def test_mps_relations():  # test/test_mps.jl:9-49, 106-135
    rng = np.random.default_rng(11)
    T = on.Tensor(rng.standard_normal((2, 2, 2)))
    mps = omps.OpenMPS(T, 3)
    assert np.allclose(oc.contract(on.Network(mps.tensors, mps.contractions, mps.openidx)),
                       omps.contract_svd_mps(mps, er=0.0), rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError, match="periodic boundary"):
        omps.contract_svd_mps(omps.PeriodicMPS(T, 3), er=0.0)
    with pytest.raises(ValueError, match="Error must be positive"):
        omps.contract_svd_mps(mps, er=-0.5)
    bs = [rng.standard_normal(2) + 1j * rng.standard_normal(2) for _ in range(5)]
    psi = bs[0]
    for b in bs[1:]:
        psi = np.kron(psi, b)
    m = omps.mps_from_vector(psi)
    assert np.allclose(oc.contract(m).reshape(-1, order="F"), psi, rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError, match="Input state must have length 2\\^N"):
        omps.mps_from_vector(np.ones(6, dtype=complex))
    mps = omps.OpenMPS(T, 3)
    mps.contractions[0] = on.Summation([(1, 2), (2, 1)])
    with pytest.raises(ValueError, match="first leg must contract with last leg"):
        omps.check_mps(mps)
    mps.contractions[0] = on.Summation([(1, 3), (2, 2)])
    with pytest.raises(ValueError, match="last leg must contract with first leg"):
        omps.check_mps(mps)


def test_switch_and_permute_vs_kron():  # test/test_mps.jl:153-170, 199-259
    rng = np.random.default_rng(12)
    N = 6
    bs = [rng.standard_normal(2) + 1j * rng.standard_normal(2) for _ in range(N)]

    def kron_all(order):
        psi = bs[order[0] - 1]
        for o in order[1:]:
            psi = np.kron(psi, bs[o - 1])
        return psi
    for order in ([2, 1, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1], [3, 1, 6, 2, 5, 4]):
        m = omps.mps_from_vector(kron_all(list(range(1, N + 1))))
        omps.permute(m, order)
        assert np.allclose(oc.contract(m).reshape(-1, order="F"), kron_all(order), rtol=1e-10, atol=1e-12)
    m = omps.mps_from_vector(kron_all(list(range(1, N + 1))))
    omps.switch(m, 2, 5)
    assert np.allclose(oc.contract(m).reshape(-1, order="F"), kron_all([1, 5, 3, 4, 2, 6]), rtol=1e-10, atol=1e-12)
    with pytest.raises(ValueError, match="must be positive"):
        omps.switch(m, 0, 2)
    with pytest.raises(ValueError, match="same length"):
        omps.permute(m, [1, 2])


def test_mpo_cnot_literal_and_apply():  # test/test_mpo.jl:62-121
    rng = np.random.default_rng(13)
    Ucnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    assert np.array_equal(og.circuit_gate(1, og.X, 2).matrix, Ucnot)
    N = 5
    psi = rng.standard_normal(2 ** N) + 1j * rng.standard_normal(2 ** N)
    for targ, cntrl in itertools.permutations(range(1, N + 1), 2):
        mps = omps.mps_from_vector(psi)
        out = oc.contract(ompo.apply_MPO(mps, ompo.MPO(Ucnot), (targ, cntrl))).reshape(-1, order="F")
        assert np.allclose(out, og.apply(psi, og.circuit_gate(targ, og.X, cntrl)), rtol=1e-10, atol=1e-12)
    N, M = 6, 3
    U = np.linalg.qr(rng.standard_normal((2 ** M, 2 ** M)) + 1j * rng.standard_normal((2 ** M, 2 ** M)))[0]
    psi = rng.standard_normal(2 ** N) + 1j * rng.standard_normal(2 ** N)
    mps = omps.mps_from_vector(psi)
    for com in itertools.combinations(range(1, N + 1), M):
        out = oc.contract(ompo.apply_MPO(mps, ompo.MPO(U), com)).reshape(-1, order="F")
        assert np.allclose(out, og.apply(psi, og.CircuitGate(com, U)), rtol=1e-10, atol=1e-12)
    for wires in ((4, 1, 6), (5, 2, 3), (2, 6, 1)):
        out = oc.contract(ompo.apply_MPO(mps, U, wires)).reshape(-1, order="F")
        assert np.allclose(out, og.apply(psi, og.CircuitGate(wires, U)), rtol=1e-10, atol=1e-12)
    assert len(ompo.MPO(U).tensors) == M
    with pytest.raises(ValueError, match="Repeated wires are not valid"):
        ompo.apply_MPO(mps, U, (1, 1, 2))


def test_sliced_contraction_equals_unsliced():  # EXTENSION: slicing invariance
    net, _, _ = ocirc.cfg2_network(10, 8, seed=3)
    o2g.optimize_contraction_order(net)
    il = oc.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    full = oc.ncon(arrays, il)
    nodes, steps = oplan.contraction_tree(il)
    dims = oplan.label_dims(arrays, il)
    S = oplan.choose_slice_labels(nodes, steps, dims, max_log2_elems=4, min_slices=8)
    assert len(S) >= 3
    assert abs(oplan.contract_sliced(arrays, il, None, S) - full) < 1e-12 * abs(full)
