"""GPU parity of the truncation path: Jacobi SVD, tail-norm rule, contract_svd, switch!,
MPS(psi), MPO and the device-resident MPS update loop, against the CPU oracle (LAPACK gesdd).
Tolerances (BASELINE.json north_star): kept singular values 1e-10, fidelity 1e-9."""
import itertools

import numpy as np
import pytest

from conftest import rel_err, to_oracle
from oracle import contract as oc
from oracle import gates as og
from oracle import mpo as ompo
from oracle import mps as omps
from oracle import mps_sim as osim
from oracle import network as on
from oracle import svd as osvd

pytestmark = pytest.mark.gpu


def sys_svd(q):
    """The ``svd`` submodule (the package attribute of that name is the function)."""
    import sys
    return sys.modules[q.__name__ + ".svd"]


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


def check_svd(q, A, er=-1.0, maxdim=0, stol=1e-10):
    U, S, Vh, k = q.svd_trunc(A, er, maxdim)
    _, Sref, _ = osvd.svd(A)
    scale = max(Sref[0], 1e-300)
    assert np.all(np.diff(S) <= 1e-14 * scale), "singular values not sorted"
    assert np.abs(S - Sref).max() <= stol * scale
    kref = len(Sref) if er < 0 else osvd.truncation_rank(Sref, er, maxdim if maxdim > 0 else None)
    if er < 0 and maxdim > 0:
        kref = min(kref, maxdim)
    assert k == kref
    rec = (U * S) @ Vh
    assert np.abs(rec - A).max() <= 1e-12 * scale * max(A.shape)
    r = len(Sref)   # both factors are isometries even when A is rank-deficient (null vectors completed, like LAPACK)
    assert np.abs(U.conj().T @ U - np.eye(r)).max() < 1e-11
    assert np.abs(Vh @ Vh.conj().T - np.eye(r)).max() < 1e-11
    return U, S, Vh, k


@pytest.mark.parametrize("shape", [(4, 4), (7, 3), (3, 7), (32, 32), (33, 17), (100, 100), (64, 200), (257, 130), (1, 5), (6, 1)])
def test_svd_shapes(gpu, shape):
    rng = np.random.default_rng(sum(shape))
    check_svd(gpu, crand(rng, *shape))


def test_svd_real_rank_deficient_and_zero(gpu):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((40, 6)) @ rng.standard_normal((6, 30)) + 0j   # rank 6
    U, S, Vh, k = check_svd(gpu, A, er=1e-9)
    assert k == 6
    U, S, Vh, k = gpu.svd_trunc(np.zeros((5, 4), dtype=complex), er=0.0)
    assert k == 0 and np.all(S == 0)                                        # reference throws; documented k = 0


def test_truncation_rule_exponential_spectrum(gpu):  # test/test_svd.jl:53-83 spectrum
    s = np.exp(-np.arange(100.0))
    rng = np.random.default_rng(2)
    qs = [np.linalg.qr(crand(rng, 100, 100))[0] for _ in range(2)]
    A = qs[0] @ np.diag(s) @ qs[1]
    U, S, Vh, k = gpu.svd_trunc(A, er=1e-10)
    assert k == osvd.truncation_rank(s, 1e-10)
    assert np.abs(S[:k] - s[:k]).max() < 1e-10
    assert np.abs(S[:12] / s[:12] - 1).max() < 1e-9    # relative accuracy down to what A itself resolves
    for er, maxdim in ((0.0, 0), (1e-3, 0), (1e-10, 5), (-1.0, 7), (3.0, 0)):
        assert gpu.svd_trunc(A, er, maxdim)[3] == min(
            osvd.truncation_rank(s, er) if er >= 0 else 100, maxdim if maxdim > 0 else 100)


def test_svd_batched_ragged(gpu):
    import ctypes as C
    from qaintensor_b200 import _lib
    rng = np.random.default_rng(3)
    shapes = [(48, 48), (20, 64), (64, 20), (130, 70), (8, 8)]
    mats = [np.asfortranarray(crand(rng, *s)) for s in shapes]
    Us = [np.zeros((m, min(m, n)), complex, order="F") for m, n in shapes]
    Ss = [np.zeros(min(m, n)) for m, n in shapes]
    Vs = [np.zeros((min(m, n), n), complex, order="F") for m, n in shapes]
    ks = (C.c_int64 * len(shapes))()
    vp = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])  # noqa: E731
    _lib.check(_lib.lib.qtn_svd_trunc_batched(len(shapes), vp(mats), _lib.arr_i64([s[0] for s in shapes]),
                                              _lib.arr_i64([s[1] for s in shapes]), 1e-12, 40, vp(Us), vp(Ss), vp(Vs), ks))
    for A, U, S, Vh, k in zip(mats, Us, Ss, Vs, ks):
        Sref = osvd.svd(A)[1]
        assert np.abs(S - Sref).max() < 1e-10 * Sref[0]
        assert k == osvd.truncation_rank(Sref, 1e-12, 40)
        assert np.abs((U * S) @ Vh - A).max() < 1e-11 * Sref[0]


def test_svd_dataflow_kernel_forced(gpu, monkeypatch):
    """The dataflow sweep kernel (jacobi_flow_kernel; default only for batches with >= 296 block pairs per round) forced on
    awkward batches (QTN_JACOBI_FLOW=2 is read per call): ragged column counts, transposed (m < n) and tall problems in
    one batch, rank-deficient and graded inputs, truncation -- against LAPACK, same bars as the three-kernel path."""
    import ctypes as C
    from qaintensor_b200 import _lib
    rng = np.random.default_rng(13)
    monkeypatch.setenv("QTN_JACOBI_FLOW", "2")
    shapes = [(200, 200), (150, 333), (333, 150), (97, 61), (256, 40), (40, 256), (129, 129)]
    mats = [np.asfortranarray(crand(rng, *s)) for s in shapes]
    mats[0][:, 100:] = mats[0][:, :100] @ crand(rng, 100, 100)                    # rank 100 of 200
    mats[6] = np.asfortranarray(mats[6] * np.exp(-np.arange(129) / 6.0)[None, :])    # graded columns (1 .. 5e-10)
    Us = [np.zeros((m, min(m, n)), complex, order="F") for m, n in shapes]
    Ss = [np.zeros(min(m, n)) for m, n in shapes]
    Vs = [np.zeros((min(m, n), n), complex, order="F") for m, n in shapes]
    ks = (C.c_int64 * len(shapes))()
    vp = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])  # noqa: E731
    before = gpu.launch_count()
    _lib.check(_lib.lib.qtn_svd_trunc_batched(len(shapes), vp(mats), _lib.arr_i64([s[0] for s in shapes]),
                                              _lib.arr_i64([s[1] for s in shapes]), 1e-9, 120, vp(Us), vp(Ss), vp(Vs), ks))
    flow_launches = gpu.launch_count() - before
    for A, U, S, Vh, k in zip(mats, Us, Ss, Vs, ks):
        Sref = osvd.svd(A)[1]
        r = len(Sref)
        assert np.abs(S - Sref).max() < 1e-10 * Sref[0]
        assert k == osvd.truncation_rank(Sref, 1e-9, 120)
        assert np.abs((U * S) @ Vh - A).max() < 1e-11 * Sref[0] * np.sqrt(r)
        assert np.abs(U.conj().T @ U - np.eye(r)).max() < 1e-11 and np.abs(Vh @ Vh.conj().T - np.eye(r)).max() < 1e-11
    # the same batch through the three-kernel path needs many more launches (3 per round)
    monkeypatch.setenv("QTN_JACOBI_FLOW", "0")
    before = gpu.launch_count()
    _lib.check(_lib.lib.qtn_svd_trunc_batched(len(shapes), vp(mats), _lib.arr_i64([s[0] for s in shapes]),
                                              _lib.arr_i64([s[1] for s in shapes]), 1e-9, 120, vp(Us), vp(Ss), vp(Vs), ks))
    assert flow_launches < (gpu.launch_count() - before) // 4
    for A, S in zip(mats, Ss):
        assert np.abs(S - osvd.svd(A)[1]).max() < 1e-10 * np.linalg.norm(A, 2)


def test_svd_1024_chi512_size(gpu):  # BASELINE config 4 SVD size, size-independent properties
    rng = np.random.default_rng(4)
    n = 1024
    s = np.concatenate([np.exp(-np.arange(600) / 40.0), np.full(n - 600, 1e-13)])
    qs = [np.linalg.qr(crand(rng, n, n))[0] for _ in range(2)]
    A = qs[0] @ np.diag(s) @ qs[1]
    U, S, Vh, k = gpu.svd_trunc(A, er=1e-10, maxdim=512)
    assert k == min(osvd.truncation_rank(np.sort(s)[::-1], 1e-10), 512)
    assert np.abs(S[:k] - np.sort(s)[::-1][:k]).max() < 1e-10
    Ak = (U[:, :k] * S[:k]) @ Vh[:k]
    assert abs(np.linalg.norm(A - Ak) - np.sqrt(np.sum(S[k:] ** 2))) < 1e-10


def test_contract_svd_relations_and_errors(gpu):  # test/test_svd.jl:7-50
    q = gpu
    rng = np.random.default_rng(5)
    t1, t2 = crand(rng, 4, 4), crand(rng, 4, 4)
    assert rel_err(q.contract_svd(q.Tensor(t1), q.Tensor(t2), (2, 1)).data, t1 @ t2) < 1e-10
    t1, t2 = crand(rng, 2, 3, 4, 6), crand(rng, 1, 5, 2, 4)
    for idx in ((1, 3), (3, 4)):
        got = q.contract_svd(q.Tensor(t1), q.Tensor(t2), idx).data
        want = osvd.contract_svd(t1, t2, idx)
        assert got.shape == want.shape and rel_err(got, want) < 1e-10
    with pytest.raises(ValueError, match="Error must be positive"):
        q.contract_svd(q.Tensor(t1), q.Tensor(t2), (1, 3), er=-0.3)
    with pytest.raises(ValueError, match="Dimensions of contraction legs do not match"):
        q.contract_svd(q.Tensor(t1), q.Tensor(crand(rng, 8, 8)), (1, 3))
    # the C ABI itself reports the reference's strings
    import ctypes as C
    from qaintensor_b200 import _lib
    a = np.asfortranarray(t1)
    rc = _lib.lib.qtn_contract_svd(a.ctypes.data, 4, _lib.arr_i64(a.shape), 1, a.ctypes.data, 4, _lib.arr_i64(a.shape), 1, -1.0, a.ctypes.data)
    assert rc == _lib.QTN_EDOMAIN and _lib.lib.qtn_last_error() == b"Error must be positive"


def test_contract_svd_truncation_bound(gpu):  # test/test_svd.jl:53-83
    q = gpu
    s = np.exp(-np.arange(100.0))
    rng = np.random.default_rng(6)
    qs = [np.linalg.qr(rng.standard_normal((100, 100)))[0] for _ in range(4)]
    T1, T2 = qs[0] @ np.diag(s) @ qs[1], qs[2] @ np.diag(s) @ qs[3]
    er = 1e-10
    mps = q.ClosedMPS([q.Tensor(T1), q.Tensor(T2)])
    approx = q.contract_svd_mps(mps, er=er)
    exact = q.contract(mps)
    assert rel_err(exact, T1 @ T2) < 1e-10
    k = osvd.truncation_rank(s, er)
    n, nt = np.linalg.norm(s), np.linalg.norm(s[k:])
    assert np.linalg.norm(approx - exact) < 2 * n * nt + nt ** 2
    assert rel_err(approx, omps.contract_svd_mps(omps.ClosedMPS([on.Tensor(T1), on.Tensor(T2)]), er=er)) < 1e-9


def test_mps_relations(gpu):  # test/test_mps.jl:9-49
    q = gpu
    rng = np.random.default_rng(7)
    T = q.Tensor(rng.standard_normal((2, 2, 2)))
    mps = q.OpenMPS(T, 3)
    g = q.GeneralTensorNetwork(mps.tensors, mps.contractions, mps.openidx)
    assert rel_err(q.contract_svd_mps(mps, er=0.0), q.contract(g)) < 1e-10
    with pytest.raises(ValueError, match="periodic boundary"):
        q.contract_svd_mps(q.PeriodicMPS(T, 3), er=0.0)
    with pytest.raises(ValueError, match="Error must be positive"):
        q.contract_svd_mps(mps, er=-0.5)
    bs = [crand(rng, 2) for _ in range(5)]
    psi = bs[0]
    for b in bs[1:]:
        psi = np.kron(psi, b)
    m = q.MPS(psi)
    assert rel_err(q.contract(m).reshape(-1, order="F"), psi) < 1e-10
    with pytest.raises(ValueError, match="Input state must have length 2\\^N"):
        q.MPS(np.ones(6, dtype=complex))
    bad = q.OpenMPS(T, 3)
    bad.contractions[0] = q.Summation([(1, 2), (2, 1)])
    with pytest.raises(ValueError, match="first leg must contract with last leg"):
        q.check_mps(bad)
    bad.contractions[0] = q.Summation([(1, 3), (2, 2)])
    with pytest.raises(ValueError, match="last leg must contract with first leg"):
        q.check_mps(bad)
    bad.tensors[1] = q.Tensor(rng.standard_normal((2, 2, 2, 2)))
    bad.contractions[0] = q.Summation([(1, 3), (2, 1)])
    with pytest.raises(ValueError, match="can only have 2 or 3 legs"):
        q.check_mps(bad)


def test_switch_and_permute_vs_kron(gpu):  # test/test_mps.jl:153-170, 199-259
    q = gpu
    rng = np.random.default_rng(8)
    N = 6
    bs = [crand(rng, 2) for _ in range(N)]

    def kron_all(order):
        psi = bs[order[0] - 1]
        for o in order[1:]:
            psi = np.kron(psi, bs[o - 1])
        return psi
    for order in ([2, 1, 3, 4, 5, 6], [3, 1, 6, 2, 5, 4]):
        m = q.MPS(kron_all(list(range(1, N + 1))))
        q.permute(m, order)
        assert rel_err(q.contract(m).reshape(-1, order="F"), kron_all(order)) < 1e-9
    m = q.MPS(kron_all(list(range(1, N + 1))))
    q.switch(m, 2, 5)
    assert rel_err(q.contract(m).reshape(-1, order="F"), kron_all([1, 5, 3, 4, 2, 6])) < 1e-9
    with pytest.raises(ValueError, match="must be positive"):
        q.switch(m, 0, 2)
    with pytest.raises(IndexError):
        q.switch(m, N)


def test_mpo_apply_vs_dense(gpu):  # test/test_mpo.jl:62-121
    q = gpu
    rng = np.random.default_rng(9)
    Ucnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    N = 5
    psi = crand(rng, 2 ** N)
    mps = q.MPS(psi)
    for targ, cntrl in ((1, 2), (4, 2), (5, 1), (2, 3)):
        out = q.contract(q.apply_MPO(mps, q.MPO(Ucnot), (targ, cntrl))).reshape(-1, order="F")
        assert rel_err(out, og.apply(psi, og.circuit_gate(targ, og.X, cntrl))) < 1e-9
    M = 3
    U = np.linalg.qr(crand(rng, 2 ** M, 2 ** M))[0]
    assert len(q.MPO(U).tensors) == M
    for wires in ((1, 3, 4), (2, 4, 5), (4, 1, 5), (5, 2, 3)):
        out = q.contract(q.apply_MPO(mps, U, wires)).reshape(-1, order="F")
        assert rel_err(out, og.apply(psi, og.CircuitGate(wires, U))) < 1e-9
        out2 = q.contract(q.apply_MPO(mps, q.CircuitGate(wires, U))).reshape(-1, order="F")
        assert rel_err(out2, out) < 1e-12
    with pytest.raises(ValueError, match="Repeated wires are not valid"):
        q.apply_MPO(mps, U, (1, 1, 2))
    with pytest.raises(ValueError, match="Direct conversion to MPS form is not support"):
        q.MPO([], [], [])


def test_decomposed_tensor_circuit(gpu):  # test/test_tensor_circuit.jl:62-126
    q = gpu
    rng = np.random.default_rng(10)
    net = q.GeneralTensorNetwork([q.Tensor(crand(rng, 2, 6)), q.Tensor(crand(rng, 2, 6, 7)), q.Tensor(crand(rng, 2, 7))],
                                 [q.Summation([(1, 2), (2, 2)]), q.Summation([(2, 3), (3, 2)])], [(1, 1), (2, 1), (3, 1)])
    cgc = [q.circuit_gate(3, q.gates.X, 1), q.circuit_gate(3, q.gates.Y, 1), q.circuit_gate(1, q.gates.Y, 2),
           q.circuit_gate(2, q.gates.Z, 1)]
    psi0 = q.contract(net).reshape(-1, order="F")
    q.tensor_circuit(net, cgc, is_decompose=True)
    ref = og.apply(psi0, [og.CircuitGate(g.iwire, g.matrix) for g in cgc])
    assert rel_err(q.contract(net).reshape(-1, order="F"), ref) < 1e-9


def test_device_mps_gate_apply_small_exact(gpu):  # EXTENSION: no truncation -> exact state
    q = gpu
    rng = np.random.default_rng(11)
    N = 8
    mps = q.DeviceMPS.product_state(N, 64)
    ref = osim.product_state(N)
    psi = np.zeros(2 ** N, complex)
    psi[0] = 1
    for layer in range(6):
        sites = q.brickwork_layer_sites(N, layer % 2)
        gates = [q.circuits.haar_unitary(4, rng) for _ in sites]
        mps.apply_layer(sites, gates, er=0.0, maxdim=0)
        osim.apply_layer(ref, sites, gates, 0.0, None)
        psi = og.apply(psi, [og.CircuitGate((s, s + 1), g) for s, g in zip(sites, gates)])
    got = osim.to_vector(mps.download())
    assert abs(abs(np.vdot(got, psi)) - 1) < 1e-10 and abs(np.linalg.norm(got) - 1) < 1e-10
    assert rel_err(got * np.vdot(got, psi), psi) < 1e-9          # equal up to the SVD's global phase freedom
    assert mps.bonds()[1] == [r.shape[2] for r in ref]
    assert abs(mps.overlap(mps) - 1) < 1e-10


def test_device_mps_truncated_fidelity_vs_oracle(gpu):  # EXTENSION: chi cap + cutoff, fidelity 1e-9
    q = gpu
    rng = np.random.default_rng(12)
    N, chi, er = 14, 16, 1e-10
    mps = q.DeviceMPS.product_state(N, chi)
    ref = osim.product_state(N)
    for layer in range(10):
        sites = q.brickwork_layer_sites(N, layer % 2)
        gates = [q.circuits.haar_unitary(4, rng) for _ in sites]
        d_gpu = mps.apply_layer(sites, gates, er=er, maxdim=chi)
        d_ref = osim.apply_layer(ref, sites, gates, er, chi)
        assert np.abs(np.array(d_gpu) - np.array(d_ref)).max() < 1e-9
    assert mps.bonds()[1] == [r.shape[2] for r in ref] and max(mps.bonds()[1]) == chi
    got = mps.download()
    nrm_g, nrm_r = osim.overlap(got, got).real, osim.overlap(ref, ref).real
    fid = abs(osim.overlap(got, ref)) ** 2 / (nrm_g * nrm_r)
    assert abs(fid - 1) < 1e-9 and abs(nrm_g - nrm_r) < 1e-9
    assert abs(mps.overlap(mps).real - nrm_g) < 1e-10


def random_open_mps(rng, bonds):
    return [crand(rng, bonds[i], 2, bonds[i + 1]) / np.sqrt(2 * bonds[i]) for i in range(len(bonds) - 1)]


def test_mpo_expectation_and_apply_exact(gpu):  # EXTENSION iii/iv, cfg 5 algorithm at an oracle-checkable size
    q = gpu
    rng = np.random.default_rng(13)
    n = 8
    sites = random_open_mps(rng, [1, 2, 4, 8, 8, 8, 4, 2, 1])
    mpo = q.tfi_mpo(n, 1.0, 0.7)
    for a, b in zip(mpo, osim.tfi_mpo(n, 1.0, 0.7)):
        assert np.array_equal(a, b)
    H = osim.mpo_to_dense(mpo)
    psi = osim.to_vector(sites)
    mps = q.DeviceMPS(sites, 64)
    want = np.vdot(psi, H @ psi)
    got = mps.expect_mpo(mpo)
    assert abs(got - want) < 1e-10 * abs(want) and abs(got - osim.expect_mpo(sites, mpo)) < 1e-10 * abs(want)
    mps.apply_mpo(mpo, er=0.0, maxdim=0)
    out = osim.to_vector(mps.download())
    ref = H @ psi
    ov = np.vdot(out, ref)
    assert abs(abs(ov) ** 2 / (np.vdot(out, out).real * np.vdot(ref, ref).real) - 1) < 1e-9
    assert rel_err(out * (ov / abs(ov)), ref) < 1e-9
    assert abs(mps.overlap(mps).real - np.vdot(ref, ref).real) < 1e-9 * np.vdot(ref, ref).real


def test_mpo_apply_compress_truncated_vs_oracle(gpu):  # fidelity 1e-9, bonds and discarded weights
    q = gpu
    rng = np.random.default_rng(14)
    n, chi = 12, 12
    bonds = [min(2 ** min(i, n - i), chi) for i in range(n + 1)]
    sites = random_open_mps(rng, bonds)
    mpo = q.tfi_mpo(n, 1.0, 1.0)
    ref = [s.copy() for s in sites]
    d_ref = osim.apply_mpo_compress(ref, mpo, 1e-10, chi)
    mps = q.DeviceMPS(sites, chi)
    d_gpu = mps.apply_mpo(mpo, er=1e-10, maxdim=chi)
    got = mps.download()
    assert [g.shape for g in got] == [r.shape for r in ref]
    assert np.abs(np.array(d_gpu) - np.array(d_ref)).max() < 1e-9 * max(1.0, max(d_ref))
    ng, nr = osim.overlap(got, got).real, osim.overlap(ref, ref).real
    assert abs(abs(osim.overlap(got, ref)) ** 2 / (ng * nr) - 1) < 1e-9 and abs(ng / nr - 1) < 1e-9
    e_gpu, e_ref = mps.expect_mpo(mpo), osim.expect_mpo(ref, mpo)
    assert abs(e_gpu - e_ref) < 1e-9 * abs(e_ref)


@pytest.mark.parametrize("shape", [(300, 200), (700, 333), (1024, 512), (128, 128)])
def test_orth_columns_cholqr2(gpu, shape):  # EXTENSION: gauge step of the MPO x MPS compression
    q = gpu
    m, n = shape
    A = crand(np.random.default_rng(m + n), m, n)
    Q, method = sys_svd(q).orth_columns(A)
    assert method == 1
    assert np.abs(Q.conj().T @ Q - np.eye(n)).max() < 1e-12
    assert np.abs(Q @ (Q.conj().T @ A) - A).max() < 1e-12 * np.abs(A).max() * n


def test_orth_columns_fallbacks(gpu):
    q = gpu
    rng = np.random.default_rng(5)
    m, n = 400, 160
    U = np.linalg.qr(crand(rng, m, n))[0]
    V = np.linalg.qr(crand(rng, n, n))[0]
    ill = (U * np.logspace(0, -13, n)) @ V.conj().T          # cond 1e13: CholeskyQR2 must refuse
    Q, method = sys_svd(q).orth_columns(ill)
    assert method == 2
    assert np.abs(Q @ (Q.conj().T @ ill) - ill).max() < 1e-12
    deficient = crand(rng, m, 40) @ crand(rng, 40, n)        # rank 40 < n
    Q, method = sys_svd(q).orth_columns(deficient)
    assert method == 2
    assert np.abs(Q @ (Q.conj().T @ deficient) - deficient).max() < 1e-11 * np.abs(deficient).max()
    small = crand(rng, 90, 60)                               # below the block size that pays
    Q, method = sys_svd(q).orth_columns(small)
    assert method == 2 and np.abs(Q.conj().T @ Q - np.eye(60)).max() < 1e-12
    moderately_ill = (U * np.logspace(0, -6, n)) @ V.conj().T  # cond 1e6: still inside CholeskyQR2's range
    Q, method = sys_svd(q).orth_columns(moderately_ill)
    assert method == 1 and np.abs(Q.conj().T @ Q - np.eye(n)).max() < 1e-12


def test_mpo_apply_compress_wide_bonds_vs_oracle(gpu):  # fat bonds >= 128: the CholeskyQR2 gauge sweep runs
    q = gpu
    rng = np.random.default_rng(21)
    n, chi = 14, 48
    bonds = [min(2 ** min(i, n - i), chi) for i in range(n + 1)]
    sites = random_open_mps(rng, bonds)
    mpo = q.tfi_mpo(n, 1.0, 0.8)
    ref = [s.copy() for s in sites]
    d_ref = osim.apply_mpo_compress(ref, mpo, 1e-10, chi)
    mps = q.DeviceMPS(sites, chi)
    d_gpu = mps.apply_mpo(mpo, er=1e-10, maxdim=chi)
    got = mps.download()
    assert [g.shape for g in got] == [r.shape for r in ref]
    assert np.abs(np.array(d_gpu) - np.array(d_ref)).max() < 1e-9 * max(1.0, max(d_ref))
    ng, nr = osim.overlap(got, got).real, osim.overlap(ref, ref).real
    assert abs(abs(osim.overlap(got, ref)) ** 2 / (ng * nr) - 1) < 1e-9 and abs(ng / nr - 1) < 1e-9
    e_gpu, e_ref = mps.expect_mpo(mpo), osim.expect_mpo(ref, mpo)
    assert abs(e_gpu - e_ref) < 1e-9 * abs(e_ref)


def inflate_bonds(rng, sites, pad):
    """Same state on bonds of size ``pad``: site_i -> W_{i-1}^H site_i W_i with random isometries W (b x pad).
    Every site matrix becomes dense and rank-deficient."""
    n = len(sites)
    out, Wl = [], np.ones((1, 1))
    for i, s in enumerate(sites):
        r = s.shape[2]
        Wr = np.ones((1, 1)) if i == n - 1 else np.linalg.qr(crand(rng, pad, r))[0].conj().T   # r x pad, W W^H = I
        out.append(np.einsum("al,lpr,rb->apb", Wl.conj().T, s, Wr))
        Wl = Wr
    return out


def test_mpo_apply_compress_rank_deficient_sites(gpu):  # every gauge-sweep matrix is rank-deficient (U-only SVD path)
    q = gpu
    rng = np.random.default_rng(33)
    n, b, pad = 10, 6, 16
    bonds = [min(2 ** min(i, n - i), b) for i in range(n + 1)]
    sites = random_open_mps(rng, bonds)
    fat = inflate_bonds(rng, sites, pad)
    assert rel_err(osim.to_vector(fat), osim.to_vector(sites)) < 1e-13
    mpo = q.tfi_mpo(n, 1.0, 0.6)
    ref = [s.copy() for s in sites]
    osim.apply_mpo_compress(ref, mpo, 1e-12, None)
    want = osim.to_vector(ref)
    mps = q.DeviceMPS(fat, 3 * pad)
    mps.apply_mpo(mpo, er=1e-12, maxdim=0)
    got = osim.to_vector(mps.download())
    assert rel_err(got, want) < 1e-9


def test_device_mps_layer_rank_deficient_sites(gpu):  # U-only gate SVDs on rank-deficient theta, er = 0 keeps noise columns
    q = gpu
    rng = np.random.default_rng(34)
    n, b, pad = 8, 4, 12
    bonds = [min(2 ** min(i, n - i), b) for i in range(n + 1)]
    sites = random_open_mps(rng, bonds)
    fat = inflate_bonds(rng, sites, pad)
    ref = [s.copy() for s in sites]
    mps = q.DeviceMPS(fat, 2 * pad)
    for layer in range(2):
        where = list(range(1 + layer % 2, n, 2))
        gates = [np.linalg.qr(crand(rng, 4, 4))[0] for _ in where]
        osim.apply_layer(ref, where, gates, 0.0, None)
        mps.apply_layer(where, gates, er=0.0, maxdim=0)
    assert rel_err(osim.to_vector(mps.download()), osim.to_vector(ref)) < 1e-10
