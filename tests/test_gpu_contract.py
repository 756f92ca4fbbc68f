"""GPU parity of the contraction path (through the C ABI) against the CPU oracle.
Tolerance: 1e-10 relative for ComplexF64 amplitudes (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import random_TN, rel_err, to_oracle
from oracle import contract as oc
from oracle import gates as og
from oracle import network2graph as o2g
from oracle import plan as oplan

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_matmul_and_two_tensor_networks(gpu):  # test/test_svd.jl:13-25 shapes
    q = gpu
    rng = np.random.default_rng(1)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    a, b = r(4, 4), r(4, 4)
    net = q.GeneralTensorNetwork([q.Tensor(a), q.Tensor(b)], [q.Summation([(1, 2), (2, 1)])], [(1, 1), (2, 2)])
    assert rel_err(q.contract(net), a @ b) < TOL
    t1, t2 = r(2, 3, 4, 6), r(1, 5, 2, 4)
    for con, opn in (([(1, 1), (2, 3)], [(1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 4)]),
                     ([(1, 3), (2, 4)], [(1, 1), (1, 2), (1, 4), (2, 1), (2, 2), (2, 3)])):
        net = q.GeneralTensorNetwork([q.Tensor(t1), q.Tensor(t2)], [q.Summation(con)], opn)
        want = oc.contract(to_oracle(net))
        got = q.contract(net)
        assert got.shape == want.shape and rel_err(got, want) < TOL


def test_single_tensor_permutedims(gpu):  # src/contract.jl:243-245
    q = gpu
    rng = np.random.default_rng(2)
    d = rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))
    net = q.GeneralTensorNetwork([q.Tensor(d)], [], [(1, 1), (1, 2)])
    assert np.array_equal(q.contract(net), d)
    net = q.GeneralTensorNetwork([q.Tensor(d)], [], [(1, 2), (1, 1)])
    assert np.array_equal(q.contract(net), d.T)


@pytest.mark.parametrize("shape,perm", [((2, 3, 4, 5), (3, 1, 4, 2)), ((7, 6), (2, 1)), ((2,) * 12, (12, 3, 1, 7, 5, 2, 9, 11, 4, 6, 8, 10)),
                                         ((64, 2, 2, 64), (1, 3, 2, 4)), ((5, 1, 3), (3, 2, 1)), ((128, 96), (2, 1)), ((4, 4, 4), (1, 2, 3))])
def test_permutedims_bit_exact(gpu, shape, perm):
    q = gpu
    rng = np.random.default_rng(3)
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    got = q.permutedims(a, perm)
    assert np.array_equal(got, np.transpose(a, [p - 1 for p in perm]))


def test_qft12_state_vector(gpu):  # BASELINE config 1
    q = gpu
    net, vecs = q.circuits.cfg1_qft_network(12)
    got = q.contract(net)
    want = oc.contract(to_oracle(net))
    assert got.shape == (2,) * 12 and rel_err(got, want) < TOL
    psi0 = vecs[0]
    for v in vecs[1:]:
        psi0 = np.kron(v, psi0)
    assert rel_err(got.reshape(-1, order="F"), og.apply(psi0, og.qft_circuit(12))) < TOL


def test_tensor_circuit_qft3_default_and_exhaustive(gpu):  # test/test_tensor_circuit.jl:32-60
    q = gpu
    rng = np.random.default_rng(4)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    net = q.GeneralTensorNetwork([q.Tensor(r(2, 6)), q.Tensor(r(2, 6, 7)), q.Tensor(r(2, 7))],
                                 [q.Summation([(1, 2), (2, 2)]), q.Summation([(2, 3), (3, 2)])],
                                 [(1, 1), (2, 1), (3, 1)])
    psi0 = q.contract(net).reshape(-1, order="F")
    assert rel_err(psi0, oc.contract(to_oracle(net)).reshape(-1, order="F")) < TOL
    q.tensor_circuit(net, q.qft_circuit(3))
    ref = og.apply(psi0, og.qft_circuit(3))
    assert rel_err(q.contract(net).reshape(-1, order="F"), ref) < TOL
    assert rel_err(q.contract(net, True).reshape(-1, order="F"), ref) < TOL


def test_random_networks_default_vs_optimized_order(gpu):  # test/test_treewidth.jl:318-347
    q = gpu
    rng = np.random.default_rng(5)
    for (Nn, Ne) in [(10, 10), (10, 20), (20, 40), (12, 30)]:
        net = random_TN(q, Nn, Ne, rng)
        want = complex(oc.contract(to_oracle(net)))
        got0 = complex(q.contract(net))
        n2 = net.copy()
        q.optimize_contraction_order(n2)
        got1 = complex(q.contract(n2))
        assert abs(got0 - want) < TOL * abs(want) and abs(got1 - want) < TOL * abs(want)


def test_self_contraction_and_disconnected(gpu):
    q = gpu
    rng = np.random.default_rng(6)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    a, b, c = r(3, 4, 3), r(4, 5), r(2, 2)
    net = q.GeneralTensorNetwork([q.Tensor(a), q.Tensor(b), q.Tensor(c)],
                                 [q.Summation([(1, 1), (1, 3)]), q.Summation([(1, 2), (2, 1)])],
                                 [(2, 2), (3, 2), (3, 1)])
    want = np.einsum("iji,jk,ab->kba", a, b, c)
    assert rel_err(q.contract(net), want) < TOL
    assert rel_err(oc.contract(to_oracle(net)), want) < 1e-12


def test_cfg2_small_and_full_amplitude(gpu):  # BASELINE config 2 at full size
    q = gpu
    for args in ((10, 8, 7), (24, 20, None)):
        net, _, _ = q.circuits.cfg2_network(*args)
        q.optimize_contraction_order(net)
        want = complex(oc.contract(to_oracle(net)))
        got = complex(q.contract(net))
        assert abs(got - want) < TOL * abs(want)


def test_plan_reuse_and_slicing_invariance(gpu):  # EXTENSION: sum over slices == unsliced
    q = gpu
    net, _, _ = q.circuits.cfg2_network(16, 12, seed=11)
    q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    shapes = [a.shape for a in arrays]
    want = complex(oc.contract(to_oracle(net)))
    plan = q.ContractionPlan(shapes, il)
    assert abs(complex(plan.execute(arrays)) - want) < TOL * abs(want)
    assert abs(complex(plan.execute(arrays)) - want) < TOL * abs(want)  # graph replay
    S = q.choose_slices(shapes, il, None, 8, 16)
    nodes, steps = oplan.contraction_tree(il)
    assert S == oplan.choose_slice_labels(nodes, steps, oplan.label_dims(arrays, il), 8, 16)
    sp = q.ContractionPlan(shapes, il, None, S)
    assert sp.nslices >= 16
    assert abs(complex(sp.execute(arrays)) - want) < TOL * abs(want)
    half = sp.nslices // 2
    parts = complex(sp.execute(arrays, 0, half)) + complex(sp.execute(None, half, sp.nslices))
    assert abs(parts - want) < TOL * abs(want)
    # per-slice parity against the oracle's tree execution
    dims = oplan.label_dims(arrays, il)
    for sid in (0, 5, sp.nslices - 1):
        ws = complex(oplan.execute_tree(arrays, il, nodes, steps, oplan.slice_assignment(S, dims, sid)))
        gs = complex(sp.execute(None, sid, sid + 1))
        assert abs(gs - ws) < TOL * max(abs(ws), abs(want))


def test_cfg3_reduced_depth_sliced(gpu):  # BASELINE config 3 topology at an oracle-sized depth
    q = gpu
    net, _, _ = q.circuits.cfg3_network(4, 4, 8, seed=21)
    q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    want = complex(oc.contract(to_oracle(net)))
    S = q.choose_slices([a.shape for a in arrays], il, None, 10, 8)
    sp = q.ContractionPlan([a.shape for a in arrays], il, None, S)
    got = complex(sp.execute(arrays))
    assert abs(got - want) < TOL * abs(want)


def test_open_legs_many_amplitudes(gpu):  # open output legs
    q = gpu
    rng = np.random.default_rng(9)
    gates = q.circuits.brickwork_gates(10, 6, rng)
    net = q.circuits.amplitude_network(10, gates, None)
    got = q.contract(net)
    psi = np.zeros(2 ** 10, dtype=complex)
    psi[0] = 1
    # non-decomposed tensor_circuit! contracts the row bits: it applies U^T (quirk Q2)
    ref = og.apply(psi, [og.CircuitGate(g.iwire, g.matrix.T) for g in gates])
    assert rel_err(got.reshape(-1, order="F"), ref) < TOL


def test_complex_f32_mode(gpu):  # EXTENSION vii: ComplexF32 mode, amplitudes within 1e-4 relative
    q = gpu
    net, _ = q.circuits.cfg1_qft_network(12)
    got = q.contract(net, precision="c64")
    want = oc.contract(to_oracle(net))
    assert got.dtype == np.complex64 and rel_err(got, want) < 1e-4
    for args in ((10, 8, 7), (24, 20, None)):
        net, _, _ = q.circuits.cfg2_network(*args)
        q.optimize_contraction_order(net)
        want = complex(oc.contract(to_oracle(net)))
        got = complex(q.contract(net, precision="c64"))
        assert abs(got - want) < 1e-4 * abs(want)
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    S = q.choose_slices([a.shape for a in arrays], il, None, 14, 8)
    sp = q.ContractionPlan([a.shape for a in arrays], il, None, S, precision="c64")
    assert abs(complex(sp.execute(arrays)) - want) < 1e-4 * abs(want)
    rng = np.random.default_rng(3)
    a = rng.standard_normal((3, 5, 4)) + 1j * rng.standard_normal((3, 5, 4))
    b = rng.standard_normal((4, 7, 5)) + 1j * rng.standard_normal((4, 7, 5))
    got = q.ncon([a, b], [[-1, 1, 2], [2, -2, 1]], precision="c64")
    assert rel_err(got, np.einsum("ijk,klj->il", a, b)) < 1e-5


def test_cfg3_full_size_parity_chain(gpu):  # BASELINE config 3 at full size (see tools/validate_cfg3.py)
    """oracle(level-28 slice) = sum of its 64 oracle level-24 sub-slices; GPU level-28 slice matches it;
    GPU level-31 slice (the default bench's tile shapes, 109 GB arena) = sum of its 32 GPU level-28 sub-slices."""
    q = gpu
    net, _, _ = q.circuits.cfg3_network()
    q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    shapes = [a.shape for a in arrays]
    S = {lvl: q.choose_slices(shapes, il, None, lvl, 1) for lvl in (24, 28, 31)}
    assert S[28][:len(S[31])] == S[31] and S[24][:len(S[28])] == S[28]
    nodes, steps = oplan.contraction_tree(il)
    dims = oplan.label_dims(arrays, il)
    n31, n28, n24 = (2 ** len(S[k]) for k in (31, 28, 24))
    sid28 = 4242
    want = sum(complex(oplan.execute_tree(arrays, il, nodes, steps, oplan.slice_assignment(S[24], dims, sid28 + n28 * t)))
               for t in range(n24 // n28))
    p28 = q.ContractionPlan(shapes, il, None, S[28])
    got = complex(p28.execute(arrays, sid28, sid28 + 1))
    assert abs(got - want) < TOL * abs(want)
    sid31 = 99
    sub = sum(complex(p28.execute(None, sid31 + n31 * t, sid31 + n31 * t + 1)) for t in range(n28 // n31))
    p28.close()
    p31 = q.ContractionPlan(shapes, il, None, S[31])
    try:
        got31 = complex(p31.execute(arrays, sid31, sid31 + 1))
    except q.QtnError as e:  # the 109 GB arena needs an otherwise idle B200
        if e.code != -4:
            raise
        pytest.skip("not enough free HBM for the 2^31-level arena: %s" % e)
    finally:
        p31.close()
    assert abs(got31 - sub) < TOL * abs(sub)


def test_complex_3m_option_matches(gpu):  # opt-in 3-multiplication complex GEMM (QTN_COMPLEX_3M=1), same 1e-10 bar
    import subprocess, sys, os
    from conftest import ROOT
    code = ("import sys; sys.path.insert(0, %r); import __graft_entry__ as g; q = g.load_package();"
            "net, _, _ = q.circuits.cfg3_network(); q.optimize_contraction_order(net); il = q.contract_rep(net);"
            "arrs = [t.data for t in net.tensors]; sh = [a.shape for a in arrs];"
            "p = q.ContractionPlan(sh, il, None, q.choose_slices(sh, il, None, 28, 1));"
            "v = complex(p.execute(arrs, 7, 8)); print(repr(v))" % ROOT)
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ, QTN_COMPLEX_3M=flag)
        outs.append(complex(eval(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1])))
    assert abs(outs[0] - outs[1]) < 1e-10 * abs(outs[0])


def test_random_general_networks_numeric(gpu):  # mixed extents, open legs, self-contractions, slicing
    from test_host_planner import _random_general_network
    q = gpu
    rng = np.random.default_rng(23)
    for trial in range(25):
        net = _random_general_network(q, rng, int(rng.integers(2, 9)), int(rng.integers(1, 12)), int(rng.integers(0, 4)))
        want = oc.contract(to_oracle(net))
        got = q.contract(net)
        assert got.shape == np.shape(want) and rel_err(got, want) < TOL
        il = q.contract_rep(net)
        arrays = [t.data for t in net.tensors]
        shapes = [a.shape for a in arrays]
        plan = q.ContractionPlan(shapes, il)
        # the target may lie below what the open legs allow: the labels found up to there still slice correctly
        S = q.choose_slices(shapes, il, None, max(int(np.log2(max(plan.max_elems, 2))) - 2, 1), 2, allow_partial=True)
        if S:
            sp = q.ContractionPlan(shapes, il, None, S)
            assert rel_err(sp.execute(arrays), want) < TOL
            assert rel_err(q.ContractionPlan(shapes, il, None, S, precision="c64").execute(arrays), want) < 1e-4


def test_contract_sliced_keyword(gpu):  # EXTENSION keyword on the reference's `contract`
    q = gpu
    net, _, _ = q.circuits.cfg2_network(14, 10, seed=2)
    q.optimize_contraction_order(net)
    want = complex(oc.contract(to_oracle(net)))
    assert abs(complex(q.contract(net, max_log2_elems=6, min_slices=8)) - want) < TOL * abs(want)
    assert abs(complex(q.contract(net, max_log2_elems=6, precision="c64")) - want) < 1e-4 * abs(want)


def test_searched_order_value_invariance(gpu):  # EXTENSION: qtn_order_search, same bar as the reference order
    q = gpu
    # (1) closed brickwork amplitude, order passed explicitly and through the optimize_contraction_order! mirror
    net, _, _ = q.circuits.cfg2_network(16, 12, seed=11)
    want = complex(oc.contract(to_oracle(net)))
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    shapes = [a.shape for a in arrays]
    order, info = q.search_order(shapes, il, 64, 3, -1)
    plan = q.ContractionPlan(shapes, il, order)
    assert plan.flops_per_slice == info["total_flops"]
    assert abs(complex(plan.execute(arrays)) - want) < TOL * abs(want)
    n2 = net.copy()
    q.optimize_contraction_order(n2, method="search", ntrials=64, seed=3)
    assert abs(complex(q.contract(n2)) - want) < TOL * abs(want)
    # (2) 2-D RQC, slicing-aware search: state-vector-like tree (huge M, small N and K), sliced
    net, _, _ = q.circuits.cfg3_network(4, 4, 8, seed=21)
    want = complex(oc.contract(to_oracle(net)))
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    shapes = [a.shape for a in arrays]
    order, info = q.search_order(shapes, il, 64, 0, 7)
    S = q.choose_slices(shapes, il, order, 7, 1)
    sp = q.ContractionPlan(shapes, il, order, S)
    assert sp.nslices == info["nslices"] and sp.max_elems <= 2 ** 7 and sp.nslices > 1
    assert abs(complex(sp.execute(arrays)) - want) < TOL * abs(want)
    assert abs(complex(q.ContractionPlan(shapes, il, order, S, precision="c64").execute(arrays)) - want) < 1e-4 * abs(want)
    # (3) open legs: 2^10 amplitudes in one contraction
    rng = np.random.default_rng(9)
    net = q.circuits.amplitude_network(10, q.circuits.brickwork_gates(10, 6, rng), None)
    want = oc.contract(to_oracle(net))
    n2 = net.copy()
    q.optimize_contraction_order(n2, method="search", ntrials=32)
    assert rel_err(q.contract(n2), want) < TOL


def test_operand_prepermute_forced(gpu):  # planner pre-permutes of scattered B operands, forced on for small networks
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prepermute_check.py")],
                         env=dict(os.environ, QTN_PREPERMUTE="2"), capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    last = out.stdout.strip().splitlines()[-1]
    assert " ok, " in last and int(last.split(" ok, ")[1].split()[0]) > 20, last


def test_sliced_self_contraction_is_the_trace(gpu):  # a sliced label with both legs on ONE tensor
    q = gpu
    rng = np.random.default_rng(31)
    r = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)  # noqa: E731
    ts = [r(3, 3, 2), r(2, 4), r(4)]
    net = q.GeneralTensorNetwork([q.Tensor(t) for t in ts],
                                 [q.Summation([(1, 1), (1, 2)]), q.Summation([(1, 3), (2, 1)]), q.Summation([(2, 2), (3, 1)])], [])
    want = complex(np.einsum("aab,bc,c->", *ts))
    il = q.contract_rep(net)
    assert abs(complex(q.contract(net)) - want) < TOL * abs(want)
    for S in ([1], [1, 2], [2, 1, 3]):   # label 1 is the self-contraction: slice d reads the diagonal T[d, d, :]
        plan = q.ContractionPlan([t.shape for t in ts], il, None, S)
        assert plan.nslices == int(np.prod([{1: 3, 2: 2, 3: 4}[l] for l in S]))
        assert abs(complex(plan.execute(ts)) - want) < TOL * abs(want)
        assert abs(complex(oplan.contract_sliced(ts, il, None, S)) - want) < 1e-12 * abs(want)


def test_stream_kernel_parity(gpu):  # persistent streaming kernel of the tall-skinny steps (M >= 16384, N <= 32, K <= 32)
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stream_check
    before = gpu.launch_count()
    assert stream_check.run() < 1e-12
    assert gpu.launch_count() > before


def test_one_shot_plan_cache_new_data_and_eviction(gpu):
    """qtn_contract keeps the plans of recently contracted network structures (api.cu plan cache): a hit must contract
    the NEW tensor data (1st use direct launches, 2nd use graph capture, 3rd+ graph replay), structures that differ
    only in one label or one extent must not hit, and more structures than cache slots must evict cleanly."""
    q = gpu
    rng = np.random.default_rng(77)
    net, _, _ = q.circuits.cfg2_network(10, 6, seed=5)
    q.optimize_contraction_order(net)
    for rep in range(5):
        for t in net.tensors:
            t.data = np.asfortranarray(rng.standard_normal(t.data.shape) + 1j * rng.standard_normal(t.data.shape))
        want = oc.contract(to_oracle(net))
        assert rel_err(q.contract(net), want) < TOL
    # same tensors, one contraction re-wired (labels differ, shapes equal): must re-plan, not hit
    a, b = net.contractions[0], net.contractions[1]
    net.contractions[0] = q.Summation([a.idx[0], b.idx[1]])
    net.contractions[1] = q.Summation([b.idx[0], a.idx[1]])
    assert rel_err(q.contract(net), oc.contract(to_oracle(net))) < TOL
    # more distinct structures than cache slots, each contracted twice (second pass hits or re-plans after eviction)
    nets = []
    seed = 100
    while len(nets) < 7:   # distinct sizes; skip draws that leave a tensor without legs (not a valid network)
        n = random_TN(q, 7 + len(nets), 16 + 2 * len(nets), np.random.default_rng(seed))
        seed += 1
        if all(t.data.shape != (1,) for t in n.tensors):
            nets.append(n)
    wants = [oc.contract(to_oracle(n)) for n in nets]
    for _ in range(2):
        for n, w in zip(nets, wants):
            assert rel_err(q.contract(n), w) < TOL
    # open legs: the cached plan reports the output shape as well
    t1 = rng.standard_normal((2, 3, 4)) + 1j * rng.standard_normal((2, 3, 4))
    t2 = rng.standard_normal((4, 5)) + 1j * rng.standard_normal((4, 5))
    netm = q.GeneralTensorNetwork([q.Tensor(t1), q.Tensor(t2)], [q.Summation([(1, 3), (2, 1)])], [(1, 1), (1, 2), (2, 2)])
    for _ in range(3):
        got = q.contract(netm)
        assert got.shape == (2, 3, 5) and rel_err(got, np.tensordot(t1, t2, axes=(2, 0))) < TOL
