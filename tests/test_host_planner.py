"""Host-side (no GPU) parity of the library's C++ planner with the oracle: contraction
order, exhaustive order, slice set and plan cost must be bit-exact."""
import numpy as np
import pytest

from conftest import random_TN, to_oracle
from oracle import contract as oc
from oracle import network2graph as o2g
from oracle import plan as oplan
from oracle.lightgraphs import complete_graph


def test_order_parity_random_networks(q):
    rng = np.random.default_rng(5)
    for (Nn, Ne) in [(10, 10), (10, 20), (10, 50), (20, 60), (30, 60), (40, 100)] * 4:
        net = random_TN(q, Nn, Ne, rng)
        try:
            want = [t[2] for t in o2g.contraction_order(to_oracle(net))]
        except StopIteration:
            continue
        perm, _ = q.network2graph.contraction_order_perm(net)
        assert perm == want


def test_order_parity_with_self_contractions(q):
    rng = np.random.default_rng(6)
    net = random_TN(q, 8, 14, rng)
    # add a self-contraction on tensor 3: two new legs joined to each other
    t = net.tensors[2]
    net.tensors[2] = q.Tensor(rng.standard_normal(t.size() + (2, 2)) + 0j)
    r = t.ndims()
    net.contractions.insert(4, q.Summation([(3, r + 1), (3, r + 2)]))
    perm, _ = q.network2graph.contraction_order_perm(net)
    assert perm == [t[2] for t in o2g.contraction_order(to_oracle(net))]
    assert perm[0] == 5


def test_hand_traced_vector(q):
    A = np.ones((2, 2), dtype=complex)
    net = q.GeneralTensorNetwork([q.Tensor(A) for _ in range(4)],
                                 [q.Summation([(1, 2), (2, 1)]), q.Summation([(2, 2), (3, 1)]),
                                  q.Summation([(3, 2), (4, 1)]), q.Summation([(4, 2), (1, 1)])], [])
    perm, tw = q.network2graph.contraction_order_perm(net)
    assert perm == [3, 4, 2, 1] and tw == 2
    assert q.contraction_order(net) == [(3, 4, 3), (1, 4, 4), (2, 3, 2), (1, 2, 1)]
    n2 = net.copy()
    q.optimize_contraction_order(n2)
    assert set(n2.contractions) == set(net.contractions) and n2.tensors == net.tensors and n2.openidx == net.openidx


def _hand_traced():
    import json
    import os
    k = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hand_traced_orders.json")))
    return {n: v for n, v in k.items() if not n.startswith("_")}


@pytest.mark.parametrize("name", sorted(_hand_traced()))
def test_hand_traced_orders_cxx(q, name):
    """qtn_order_treewidth against the hand traces of tests/golden/hand_traced_orders.json."""
    k = _hand_traced()[name]
    net = q.GeneralTensorNetwork([q.Tensor(np.ones((2,) * l, dtype=complex)) for l in k["legs"]],
                                 [q.Summation([tuple(p) for p in c]) for c in k["contractions"]], [])
    perm, tw = q.network2graph.contraction_order_perm(net)
    assert perm == k["perm"] and tw == k["tw"]
    assert [list(t) for t in q.contraction_order(net)] == k["order"]


def test_treewidth_known_answers_cxx(q):  # test/test_treewidth.jl:204-221 through the C ABI
    for n in (10, 25, 50):
        tw, order = q.tree_decomposition_width(n, complete_graph(n).edges())
        assert tw == n - 1 and sorted(order) == list(range(1, n + 1))
    for k in range(2, 6):
        G = o2g.local_circuit_graph(10, k)
        tw, order = q.tree_decomposition_width(10, G.edges())
        assert tw == k - 1 and order == o2g.min_fill_ordering(G)


def test_minfill_order_parity_random_graphs(q):
    rng = np.random.default_rng(7)
    for _ in range(10):
        Nn = int(rng.integers(8, 40))
        Ne = int(rng.integers(Nn, Nn * (Nn - 1) // 2 + 1))
        G = o2g.random_graph(Nn, Ne, rng)
        tw, order = q.tree_decomposition_width(Nn, G.edges())
        assert order == o2g.min_fill_ordering(G) and tw == o2g.tree_decomposition(G)[0]


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_baseline_configs_order_slices_cost(q, name):
    net = (q.circuits.cfg2_network() if name == "cfg2" else q.circuits.cfg3_network())[0]
    onet = to_oracle(net)
    want = o2g.optimize_contraction_order(onet)
    perm, tw = q.network2graph.contraction_order_perm(net)
    assert perm == want and tw == (26 if name == "cfg2" else 54)
    q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    assert il == oc.contract_rep(onet)
    shapes = [t.size() for t in net.tensors]
    plan = q.ContractionPlan(shapes, il)
    nodes, steps = oplan.contraction_tree(il)
    dims = oplan.label_dims([t.data for t in net.tensors], il)
    f, b, mx, mnk = oplan.tree_cost(nodes, steps, dims)
    assert (plan.n_pairwise(), plan.flops_per_slice, plan.bytes_per_slice, plan.max_elems) == (len(steps), f, b, mx)
    got = sorted((max(m, n), min(m, n), k) for m, n, k, fl in plan.steps() if (fl >> 1) & 7 == 0)  # pairwise steps only
    assert got == sorted((max(m, n), min(m, n), k) for m, n, k in mnk)
    for lim, mins in ((28, 1), (30, 1)) if name == "cfg3" else ((16, 1), (12, 64)):
        S = q.choose_slices(shapes, il, None, lim, mins)
        assert S == oplan.choose_slice_labels(nodes, steps, dims, lim, mins)
        sp = q.ContractionPlan(shapes, il, None, S)
        f2, b2, mx2, _ = oplan.tree_cost(nodes, steps, dims, S)
        assert (sp.flops_per_slice, sp.bytes_per_slice, sp.max_elems) == (f2, b2, mx2)
        assert sp.nslices == 2 ** len(S) and mx2 <= 2 ** lim


def test_exhaustive_order_parity(q):  # src/contract.jl:184-235 incl. quirk Q1 inputs
    from oracle import gates as og, mpo as ompo, network as on
    rng = np.random.default_rng(8)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    ts = [r(2, 6), r(2, 6, 7), r(2, 7)]
    net = q.GeneralTensorNetwork([q.Tensor(t) for t in ts],
                                 [q.Summation([(1, 2), (2, 2)]), q.Summation([(2, 3), (3, 2)])],
                                 [(1, 1), (2, 1), (3, 1)])
    q.tensor_circuit(net, q.qft_circuit(3))
    leg_costs, il = q.contract_rep(net, True)
    seq, cost = q.contract_order(net, leg_costs, il)
    onet = to_oracle(net)
    lc, oil = oc.contract_rep(onet, True)
    oseq, ocost = oc.contract_order(onet, lc, oil)
    assert (seq, cost) == (oseq, ocost)
    for trial in range(4):
        net = random_TN(q, 6, 9, rng)
        if any(t.size() == (1,) for t in net.tensors):
            continue
        leg_costs, il = q.contract_rep(net, True)
        onet = to_oracle(net)
        try:
            want = oc.contract_order(onet, *oc.contract_rep(onet, True))
        except Exception:
            continue
        assert q.contract_order(net, leg_costs, il) == want


def test_plan_rejects_malformed_networks(q):
    with pytest.raises(q.QtnError, match="appears 1 times"):
        q.ContractionPlan([(2, 2), (2,)], [[1, 2], [1]])
    with pytest.raises(q.QtnError, match="different extent"):
        q.ContractionPlan([(2, 3), (2,)], [[-1, 1], [1]])
    with pytest.raises(q.QtnError, match="not a contracted label"):
        q.ContractionPlan([(2, 2), (2,)], [[-1, 1], [1]], None, [7])


def _random_general_network(q, rng, nt, ncon, nopen):
    """Random network with mixed extents (2..4), open legs and an occasional self-contraction."""
    legs = [[] for _ in range(nt)]          # per tensor: list of extents
    cons = []
    for _ in range(ncon):
        d = int(rng.integers(2, 5))
        if rng.random() < 0.1:
            a = b = int(rng.integers(1, nt + 1))
        else:
            a, b = (int(x) for x in rng.choice(np.arange(1, nt + 1), size=2, replace=False))
        legs[a - 1].append(d)
        la = len(legs[a - 1])
        legs[b - 1].append(d)
        cons.append(q.Summation([(a, la), (b, len(legs[b - 1]))]))
    opn = []
    for _ in range(nopen):
        t = int(rng.integers(1, nt + 1))
        legs[t - 1].append(int(rng.integers(2, 4)))
        opn.append((t, len(legs[t - 1])))
    for t in range(nt):
        if not legs[t]:
            legs[t].append(2)
            opn.append((t + 1, 1))
    ts = [q.Tensor(rng.standard_normal(tuple(l)) + 1j * rng.standard_normal(tuple(l))) for l in legs]
    return q.GeneralTensorNetwork(ts, cons, opn)


def test_random_general_networks_plan_parity(q):
    """Planner vs oracle tree on random networks: step count, (M,N,K) multiset, cost, output shape, slice set."""
    rng = np.random.default_rng(11)
    for trial in range(40):
        nt = int(rng.integers(2, 9))
        net = _random_general_network(q, rng, nt, int(rng.integers(1, 12)), int(rng.integers(0, 4)))
        il = q.contract_rep(net)
        arrays = [t.data for t in net.tensors]
        shapes = [a.shape for a in arrays]
        plan = q.ContractionPlan(shapes, il)
        nodes, steps = oplan.contraction_tree(il)
        dims = oplan.label_dims(arrays, il)
        f, b, mx, mnk = oplan.tree_cost(nodes, steps, dims)
        gemm = [(max(m, n), min(m, n), k) for m, n, k, fl in plan.steps() if ((fl >> 1) & 7) == 0]
        assert sorted(gemm) == sorted((max(m, n), min(m, n), k) for m, n, k in mnk)
        assert (plan.flops_per_slice, plan.bytes_per_slice) == (f, b)
        want_shape = tuple(dims[l] for l in sorted((l for lab in il for l in lab if l < 0), reverse=True))
        assert plan.out_dims == want_shape
        lim = max(int(np.log2(max(mx, 2))) - 2, 1)
        S = q.choose_slices(shapes, il, None, lim, 2, allow_partial=True)
        assert S == oplan.choose_slice_labels(nodes, steps, dims, lim, 2)
        _, _, mx_after, _ = oplan.tree_cost(nodes, steps, dims, S)
        nsl = int(np.prod([dims[l] for l in S])) if S else 1
        if mx_after > 2 ** lim or nsl < 2:   # an unreachable target is reported, not silently ignored
            with pytest.raises(q.QtnError, match="qtn_choose_slices"):
                q.choose_slices(shapes, il, None, lim, 2)
        else:
            assert S == q.choose_slices(shapes, il, None, lim, 2)
        if S:
            sp = q.ContractionPlan(shapes, il, None, S)
            f2, b2, mx2, _ = oplan.tree_cost(nodes, steps, dims, S)
            assert (sp.flops_per_slice, sp.nslices) == (f2, int(np.prod([dims[l] for l in S])))


# ---- EXTENSION: qtn_order_search ---------------------------------------------------
def _plan_total(q, shapes, il, order, max_log2):
    S = q.choose_slices(shapes, il, order, max_log2, 1) if max_log2 >= 0 else []
    plan = q.ContractionPlan(shapes, il, order, S)
    tot = plan.flops_per_slice * plan.nslices
    info = (tot, plan.flops_per_slice, plan.nslices, plan.max_elems)
    plan.close()
    return info


def test_order_search_is_valid_deterministic_and_exactly_costed(q):
    net, _, _ = q.circuits.cfg2_network(12, 8, seed=3)
    il = q.contract_rep(net)
    shapes = [t.data.shape for t in net.tensors]
    o1, i1 = q.search_order(shapes, il, 64, 7, -1)
    o2, i2 = q.search_order(shapes, il, 64, 7, -1)
    assert o1 == o2 and i1 == i2                                   # deterministic for (network, ntrials, seed)
    assert sorted(o1) == list(range(1, len(net.contractions) + 1))  # a complete sequence of the contracted labels
    tot, fps, nsl, mx = _plan_total(q, shapes, il, o1, -1)
    assert i1["total_flops"] == tot and i1["flops_per_slice"] == fps and i1["nslices"] == nsl == 1
    assert 2.0 ** i1["log2_max_elems"] == mx                       # the reported cost IS the planner's cost of that order
    o3, _ = q.search_order(shapes, il, 64, 8, -1)
    assert sorted(o3) == sorted(o1)


def test_order_search_beats_reference_order_on_cfg3_and_cfg2(q):
    # cfg 3 (6x6, 16 cycles): reference treewidth order 3.2e17 flop at the 2^31 slicing; the search must find < 1e15
    net, _, _ = q.circuits.cfg3_network()
    il = q.contract_rep(net)
    shapes = [t.data.shape for t in net.tensors]
    order, info = q.search_order(shapes, il, 128, 0, 31)
    tot, _, nsl, mx = _plan_total(q, shapes, il, order, 31)
    assert tot == info["total_flops"] and nsl == info["nslices"] and mx <= 2 ** 31
    # independent check: the oracle's walk, cost model and slice rule applied to the searched label sequence
    nodes, steps = oplan.contraction_tree(il, order)
    dims = oplan.label_dims([t.data for t in net.tensors], il)
    S = oplan.choose_slice_labels(nodes, steps, dims, 31, 1)
    assert S == q.choose_slices(shapes, il, order, 31, 1)
    f, _, mx_o, _ = oplan.tree_cost(nodes, steps, dims, S)
    assert f == info["flops_per_slice"] and 2 ** len(S) == nsl and mx_o == mx
    ref = net.copy()
    q.optimize_contraction_order(ref)
    ref_tot = _plan_total(q, shapes, q.contract_rep(ref), None, 31)[0]
    assert ref_tot > 3e17 and tot < 1e15
    net2, _, _ = q.circuits.cfg2_network()
    il2 = q.contract_rep(net2)
    shapes2 = [t.data.shape for t in net2.tensors]
    _, info2 = q.search_order(shapes2, il2, 256, 0, -1)
    ref2 = net2.copy()
    q.optimize_contraction_order(ref2)
    assert info2["total_flops"] < _plan_total(q, shapes2, q.contract_rep(ref2), None, -1)[0]


def test_order_search_general_networks_and_mirror(q):
    rng = np.random.default_rng(11)
    r = lambda *sh: rng.standard_normal(sh) + 1j * rng.standard_normal(sh)  # noqa: E731
    disconnected = q.GeneralTensorNetwork(
        [q.Tensor(r(2, 3)), q.Tensor(r(3, 2)), q.Tensor(r(4, 2)), q.Tensor(r(2, 4)), q.Tensor(r(2, 2))],
        [q.Summation([(1, 1), (2, 2)]), q.Summation([(1, 2), (2, 1)]), q.Summation([(3, 1), (4, 2)]),
         q.Summation([(3, 2), (4, 1)]), q.Summation([(5, 1), (5, 2)])], [])
    nets = [random_TN(q, Nn, Ne, rng) for (Nn, Ne) in [(2, 1), (3, 6), (6, 12), (10, 30), (20, 60)]]
    nets = [n for n in nets if all(t.data.shape != (1,) for t in n.tensors)] + [disconnected]
    assert len(nets) >= 4
    for net in nets:
        Ne = len(net.contractions)
        il = q.contract_rep(net)
        shapes = [t.data.shape for t in net.tensors]
        order, info = q.search_order(shapes, il, 16, 1, -1)
        assert sorted(order) == list(range(1, Ne + 1))
        assert info["total_flops"] == _plan_total(q, shapes, il, order, -1)[0]
        # value invariance under the searched order (test/test_treewidth.jl:326-327 for the reference's order)
        n2 = net.copy()
        q.optimize_contraction_order(n2, method="search", ntrials=16, seed=1)
        assert set(n2.contractions) == set(net.contractions) and n2.tensors == net.tensors
        assert [net.contractions[k - 1] for k in order] == n2.contractions
        assert np.allclose(oc.contract(to_oracle(n2)), oc.contract(to_oracle(net)), rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        q.optimize_contraction_order(net, method="nope")
    with pytest.raises(q.QtnError):
        q.search_order([(2, 2), (2, 2)], [[1, 2], [2, 1]], 0, 0, -1)  # ntrials < 1


def test_order_search_random_general_networks_vs_oracle(q):
    """Searched orders on random general networks (mixed extents, open legs, self-contractions, memory targets):
    complete label sequence, reported cost == planner cost == oracle cost of that sequence, oracle slice rule agrees."""
    rng = np.random.default_rng(17)
    for trial in range(30):
        nt = int(rng.integers(2, 10))
        net = _random_general_network(q, rng, nt, int(rng.integers(1, 14)), int(rng.integers(0, 4)))
        il = q.contract_rep(net)
        arrays = [t.data for t in net.tensors]
        shapes = [a.shape for a in arrays]
        dims = oplan.label_dims(arrays, il)
        ncon = len(net.contractions)
        order, info = q.search_order(shapes, il, 24, trial, -1)
        assert sorted(order) == list(range(1, ncon + 1))
        nodes, steps = oplan.contraction_tree(il, order)
        f, b, mx, _ = oplan.tree_cost(nodes, steps, dims)
        assert (info["total_flops"], info["nslices"], round(2.0 ** info["log2_max_elems"])) == (f, 1.0, mx)
        plan = q.ContractionPlan(shapes, il, order)
        assert (plan.flops_per_slice, plan.bytes_per_slice) == (f, b)
        # with a memory target: the slice set of the returned order under the deterministic rule, oracle side
        lim = max(int(np.log2(max(mx, 2))) - 2, 1)
        order2, info2 = q.search_order(shapes, il, 24, trial, lim)
        nodes2, steps2 = oplan.contraction_tree(il, order2)
        S = oplan.choose_slice_labels(nodes2, steps2, dims, lim, 1)
        assert S == q.choose_slices(shapes, il, order2, lim, 1, allow_partial=True)   # open legs may keep a tensor above the target
        f2, _, _, _ = oplan.tree_cost(nodes2, steps2, dims, S)
        nsl = int(np.prod([dims[l] for l in S])) if S else 1
        assert (info2["flops_per_slice"], info2["nslices"]) == (f2, float(nsl))


def test_split_k_rules_of_latency_bound_plans(q):
    """Kernel selection of the planner (device tuning, not part of the reference's result): split-K only where the tile
    grid cannot fill the 148 SMs -- K >= 256 steps, and small-output (<= 64 K elements) K >= 32 steps down to one 8-deep
    k-iteration per CTA; flops / bytes / step shapes of the plan are unaffected (they are compared with the oracle above)."""
    net, _, _ = q.circuits.cfg2_network()
    q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    plan = q.ContractionPlan([t.data.shape for t in net.tensors], il, None, [])
    steps = plan.steps()
    by_shape = {(m, n, k): fl >> 8 for m, n, k, fl in steps if (fl >> 1) & 7 == 0}
    assert by_shape[(256, 128, 128)] == 16          # 8 tiles, K = 128 -> one k-iteration per CTA
    assert by_shape[(256, 256, 2048)] == 19         # 16 tiles -> 2 * 148 / 16
    assert by_shape[(4096, 128, 128)] == 1          # 128 tiles and a 512 K-element output: atomics would cost more
    assert by_shape[(1024, 512, 32)] == 1
    for m, n, k, fl in steps:
        split = fl >> 8
        assert split >= 1 and (split == 1 or k // split >= 8)
    plan.close()
