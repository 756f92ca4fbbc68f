"""GPU parity at BASELINE sizes: the Jacobi path inside the MPS / MPO loops at
chi = 256 / 512 against the CPU oracle (numpy GEMM + LAPACK zgesdd), which follows
src/switch.jl:26-55 (two-site theta -> svd -> U / S*V' split) and src/svd.jl:29-33 (tail-norm rule)
plus the chi cap.  Tolerances (BASELINE.json north_star): kept singular values 1e-10 (of sigma_max),
truncated-state fidelity 1e-9, bond dimensions and kept counts exact."""
import numpy as np
import pytest

from oracle import mps_sim as osim
from oracle import svd as osvd

pytestmark = pytest.mark.gpu


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


def saturated_mps(nsites, chi, rng):
    """Random MPS with the saturated bond profile min(2^i, 2^(N-i), chi) (the steady state of cfg 4 / cfg 5)."""
    bonds = [min(2 ** min(i, nsites - i), chi) for i in range(nsites + 1)]
    return [crand(rng, bonds[i], 2, bonds[i + 1]) / np.sqrt(2.0 * bonds[i]) for i in range(nsites)]


def kept_sigmas(site):
    """sigma_j of the bond to the left of `site` = S*V' (k, 2, R): the rows of V' are orthonormal."""
    return np.linalg.norm(np.reshape(site, (site.shape[0], -1)), axis=1)


def fidelity(a, b):
    return abs(osim.overlap(a, b)) ** 2 / (osim.overlap(a, a).real * osim.overlap(b, b).real)


def test_cfg4_half_layer_chi256_vs_oracle(gpu):
    """One brickwork half-layer at chi = 256 (theta up to 512 x 512, ragged batch incl. the small edge bonds)."""
    q = gpu
    rng = np.random.default_rng(20261017 + 4256)
    N, chi, er = 22, 256, 1e-10
    sites = saturated_mps(N, chi, rng)
    mps = q.DeviceMPS(sites, chi)
    ref = [s.copy() for s in sites]
    left = q.brickwork_layer_sites(N, 1)          # bonds 2, 4, ..., 20: five of them saturated at 256
    gates = [q.circuits.haar_unitary(4, rng) for _ in left]
    d_gpu = mps.apply_layer(left, gates, er=er, maxdim=chi)
    d_ref = osim.apply_layer(ref, left, gates, er, chi)
    got = mps.download()
    assert mps.bonds()[1] == [r.shape[2] for r in ref] and max(mps.bonds()[1]) == chi
    scale = max(kept_sigmas(r).max() for r in ref)
    for s in left:  # kept singular values of every updated bond
        sg, sr = kept_sigmas(got[s]), kept_sigmas(ref[s])
        assert sg.shape == sr.shape and np.abs(sg - sr).max() <= 1e-10 * scale
    assert np.abs(np.array(d_gpu) - np.array(d_ref)).max() <= 1e-10 * scale
    assert abs(fidelity(got, ref) - 1) < 1e-9
    for s in left:  # left sites are isometries (U of the SVD)
        A = np.reshape(got[s - 1], (-1, got[s - 1].shape[2]), order="F")
        assert np.abs(A.conj().T @ A - np.eye(A.shape[1])).max() < 1e-11
    mps.close()


def test_cfg4_bond_chi512_vs_oracle(gpu):
    """One gate on a saturated chi = 512 bond: theta is 1024 x 1024 (the BASELINE config-4 SVD), cut back to 512."""
    q = gpu
    rng = np.random.default_rng(20261017 + 4512)
    N, chi, er = 20, 512, 1e-10
    sites = saturated_mps(N, chi, rng)
    mps = q.DeviceMPS(sites, chi)
    ref = [s.copy() for s in sites]
    gate = q.circuits.haar_unitary(4, rng)
    s = 10                                        # bonds 9, 10, 11 are all 512
    assert sites[s - 1].shape == (512, 2, 512) and sites[s].shape == (512, 2, 512)
    d_gpu = mps.apply_gate2(s, gate, er=er, maxdim=chi)
    d_ref = osim.apply_gate2(ref, s, gate, er, chi)
    got = mps.download()
    assert got[s].shape == ref[s].shape == (512, 2, 512)
    sg, sr = kept_sigmas(got[s]), kept_sigmas(ref[s])
    assert np.abs(sg - sr).max() <= 1e-10 * sr.max()
    assert abs(d_gpu - d_ref) <= 1e-10 * sr.max()
    # truncated two-site block: U_k S_k V_k' is unique although U, V are not
    tg = np.reshape(got[s - 1], (1024, 512), order="F") @ np.reshape(got[s], (512, 1024), order="F")
    tr = np.reshape(ref[s - 1], (1024, 512), order="F") @ np.reshape(ref[s], (512, 1024), order="F")
    assert np.linalg.norm(tg - tr) <= 1e-9 * np.linalg.norm(tr)
    assert abs(fidelity(got, ref) - 1) < 1e-9
    mps.close()


def test_cfg5_site_compress_3072x1536_vs_lapack(gpu):
    """The (2*chi*D x chi*D) = 3072 x 1536 fat site of the N = 40, chi = 512, D = 3 profile: MPO site applied to a random
    MPS site, then (a) the gauge step (orthonormal basis + carry), (b) its truncated SVD against LAPACK."""
    import sys
    q = gpu
    svdmod = sys.modules[q.__name__ + ".svd"]
    rng = np.random.default_rng(20261017 + 5512)
    chi, D = 512, 3
    A = crand(rng, chi, 2, chi) / np.sqrt(2.0 * chi)
    W = np.asarray(q.tfi_mpo(5, 1.0, 1.0)[2])                      # interior site (3, 2, 2, 3)
    B = np.einsum("aqpb,lpr->laqrb", W, A)
    M = np.asfortranarray(np.reshape(B, (chi * D * 2, chi * D), order="F"))
    # a TFI site applied to a random tensor is rank-deficient by construction? no: (l a q) x (r b) has full column rank
    # only if the three operator strings are independent on this site -- check what LAPACK sees and test against it
    Sref = osvd.svd(M)[1]
    U, S, Vh, k = q.svd_trunc(M, er=1e-10, maxdim=chi)
    assert np.abs(S - Sref).max() <= 1e-10 * Sref[0]
    assert k == min(osvd.truncation_rank(Sref, 1e-10), chi)
    Mk = (U[:, :k] * S[:k]) @ Vh[:k]
    assert abs(np.linalg.norm(M - Mk) - np.sqrt(np.sum(Sref[k:] ** 2))) <= 1e-10 * Sref[0]
    assert np.abs(U[:, :k].conj().T @ U[:, :k] - np.eye(k)).max() < 1e-11
    Q, method = svdmod.orth_columns(M)
    assert np.abs(Q.conj().T @ Q - np.eye(chi * D)).max() < 1e-11
    assert np.abs(Q @ (Q.conj().T @ M) - M).max() <= 1e-11 * np.abs(M).max() * np.sqrt(chi * D)


def test_cfg5_apply_mpo_chi512_vs_oracle(gpu):
    """MPO x MPS apply + compress at chi = 512 on a 20-site chain (bonds 9..11 saturated: the 3072 x 1536 gauge SVDs and
    the 1536 x 1024 truncating SVDs of cfg 5) against the oracle's LAPACK sweeps: bonds, discarded weights, energy,
    fidelity."""
    q = gpu
    rng = np.random.default_rng(20261017 + 5020)
    N, chi, er = 20, 512, 1e-10
    sites = saturated_mps(N, chi, rng)
    nrm = np.sqrt(osim.overlap(sites, sites).real)
    sites[N // 2] = sites[N // 2] / nrm
    mpo = q.tfi_mpo(N, 1.0, 1.0)
    mps = q.DeviceMPS(sites, chi)
    ref = [s.copy() for s in sites]
    d_gpu = mps.apply_mpo(mpo, er=er, maxdim=chi)
    d_ref = osim.apply_mpo_compress(ref, mpo, er, chi)
    got = mps.download()
    assert mps.bonds()[1] == [r.shape[2] for r in ref] and max(mps.bonds()[1]) == chi
    nr = np.sqrt(osim.overlap(ref, ref).real)
    assert np.abs(np.array(d_gpu) - np.array(d_ref)).max() <= 1e-9 * nr
    assert abs(fidelity(got, ref) - 1) < 1e-9
    assert abs(osim.overlap(got, got).real - nr ** 2) <= 1e-9 * nr ** 2
    e_gpu, e_ref = mps.expect_mpo(mpo), osim.expect_mpo(ref, mpo)
    assert abs(e_gpu - e_ref) <= 1e-9 * abs(e_ref)
    mps.close()


def test_mps_from_vector_18_sites(gpu):
    """MPS(psi) for M >= 17 -- the first SVD is 2 x 2^(M-1); only min(m, n) is bounded now
    (src/mps.jl:55-89 handles any M through LAPACK)."""
    q = gpu
    rng = np.random.default_rng(18)
    M = 18
    psi = crand(rng, 2 ** M)
    psi /= np.linalg.norm(psi)
    m = q.MPS(psi)
    assert len(m.tensors) == M and m.tensors[8].data.shape == (256, 2, 512)
    out = q.contract(m).reshape(-1, order="F")
    assert np.abs(out - psi).max() <= 1e-10 * np.abs(psi).max()
    # left-canonical form: every site but the last is an isometry (U of its SVD)
    for t in m.tensors[:-1]:
        A = np.reshape(t.data, (-1, t.data.shape[-1]), order="F")
        assert np.abs(A.conj().T @ A - np.eye(A.shape[1])).max() < 1e-11
