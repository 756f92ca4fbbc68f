"""Oracle: MPS gate application with truncation (test infrastructure, EXTENSION).

The reference has no two-site gate update; it is composed from reference primitives as
SURVEY.md section 3.5 lays out: theta = T_i * T_{i+1} (src/switch.jl:26 with er = 0) ->
4x4 gate on the two physical legs -> ``svd`` (src/switch.jl:39) -> truncate with the rule
of src/svd.jl:29-33 plus ``k <- min(k, maxdim)`` -> ``U`` / ``diagm(S)*V'`` split
(src/switch.jl:50-52).  Site layout (lbond, 2, rbond) (src/mps.jl:99-110); the gate matrix
index is p_i + 2*p_{i+1}.
"""
import numpy as np

from .svd import svd, truncation_rank


def product_state(n, vectors=None):
    out = []
    for i in range(n):
        v = np.array([1.0, 0.0], dtype=np.complex128) if vectors is None else np.asarray(vectors[i], np.complex128)
        out.append(v.reshape(1, 2, 1))
    return out


def apply_gate2(sites, i, gate, er=0.0, maxdim=None):
    """In place on the list ``sites``; ``i`` is the 1-based left site.  Returns the discarded weight."""
    A, B = sites[i - 1], sites[i]
    L, _, b = A.shape
    _, _, R = B.shape
    theta = np.reshape(A, (2 * L, b), order="F") @ np.reshape(B, (b, 2 * R), order="F")
    t4 = np.reshape(theta, (L, 2, 2, R), order="F")
    G = np.reshape(np.asarray(gate, np.complex128), (2, 2, 2, 2), order="F")  # (p1', p2', p1, p2)
    t4 = np.einsum("abcd,lcdr->labr", G, t4)
    U, S, Vh = svd(np.reshape(t4, (2 * L, 2 * R), order="F"))
    k = max(truncation_rank(S, er, maxdim), 1)
    sites[i - 1] = np.reshape(U[:, :k], (L, 2, k), order="F")
    sites[i] = np.reshape(S[:k, None] * Vh[:k, :], (k, 2, R), order="F")
    return float(np.sqrt(np.sum(S[k:] ** 2)))


def apply_layer(sites, left_sites, gates, er=0.0, maxdim=None):
    return [apply_gate2(sites, s, g, er, maxdim) for s, g in zip(left_sites, gates)]


def overlap(a, b):
    """<a|b> by transfer matrices."""
    E = np.ones((1, 1), dtype=np.complex128)
    for A, B in zip(a, b):
        T = np.tensordot(E, B, axes=(1, 0))                      # (x, p, b): two GEMM-shaped steps, O(chi^3)
        E = np.tensordot(np.conj(A), T, axes=((0, 1), (0, 1)))   # (a, b)
    return complex(E[0, 0])


def to_vector(sites):
    """Dense state, site 1 = fastest-varying bit (only for small N)."""
    psi = np.ones((1, 1), dtype=np.complex128)  # (phys..., bond)
    for A in sites:
        psi = np.tensordot(psi, A, axes=(psi.ndim - 1, 0))
    psi = psi.reshape(psi.shape[1:-1])
    return np.reshape(psi, (-1,), order="F")


# ---- MPO x MPS (EXTENSION iii / iv of SURVEY section 8a) --------------------------------------------
def tfi_mpo(n, J=1.0, h=1.0):
    """H = -J sum Z_i Z_{i+1} - h sum X_i as a D = 3 site-tensor MPO, layout (bond_in, out, in,
    bond_out) of src/mpo.jl:66; boundary bonds 1."""
    I2 = np.eye(2)
    X = np.array([[0, 1], [1, 0]], dtype=float)
    Z = np.diag([1.0, -1.0])
    Wm = np.zeros((3, 2, 2, 3), dtype=np.complex128)
    Wm[0, :, :, 0] = I2
    Wm[1, :, :, 0] = Z
    Wm[2, :, :, 0] = -h * X
    Wm[2, :, :, 1] = -J * Z
    Wm[2, :, :, 2] = I2
    sites = []
    for i in range(n):
        w = Wm
        if i == 0:
            w = w[2:3]          # start in the "nothing applied yet" row
        if i == n - 1:
            w = w[:, :, :, 0:1]  # end in the "everything applied" column
        sites.append(np.ascontiguousarray(w))
    return sites


def mpo_to_dense(mpo):
    """Dense 2^n x 2^n matrix, site 1 = fastest bit (small n only)."""
    n = len(mpo)
    T = mpo[0][0]  # (out, in, bond)
    for w in mpo[1:]:
        T = np.tensordot(T, w, axes=(T.ndim - 1, 0))
    T = T[..., 0]
    outs = [2 * i for i in range(n)]
    ins = [2 * i + 1 for i in range(n)]
    T = np.transpose(T, outs + ins)
    return np.reshape(T, (2 ** n, 2 ** n), order="F")


def apply_mpo_compress(sites, mpo, er=0.0, maxdim=None):
    """In place: site-wise apply, left-to-right SVD orthogonalisation, right-to-left truncating sweep."""
    n = len(sites)
    fat = []
    for A, W in zip(sites, mpo):
        L, _, R = A.shape
        Dl, _, _, Dr = W.shape
        B = np.einsum("aqpb,lpr->laqrb", W, A)            # (l, a, p', r, b)
        fat.append(np.reshape(B, (L * Dl, 2, R * Dr), order="F"))
    for i in range(n - 1):
        l, _, r = fat[i].shape
        U, S, Vh = svd(np.reshape(fat[i], (2 * l, r), order="F"))
        k = len(S)
        fat[i] = np.reshape(U, (l, 2, k), order="F")
        nxt = fat[i + 1]
        fat[i + 1] = np.reshape((S[:, None] * Vh) @ np.reshape(nxt, (r, -1), order="F"), (k, 2, nxt.shape[2]), order="F")
    disc = [0.0] * (n - 1)
    for i in range(n - 1, 0, -1):
        l, _, r = fat[i].shape
        U, S, Vh = svd(np.reshape(fat[i], (l, 2 * r), order="F"))
        k = max(truncation_rank(S, er, maxdim), 1)
        disc[i - 1] = float(np.sqrt(np.sum(S[k:] ** 2)))
        fat[i] = np.reshape(Vh[:k], (k, 2, r), order="F")
        prv = fat[i - 1]
        fat[i - 1] = np.reshape(np.reshape(prv, (-1, l), order="F") @ (U[:, :k] * S[:k]), (prv.shape[0], 2, k), order="F")
    sites[:] = fat
    return disc


def expect_mpo(sites, mpo):
    """<psi| MPO |psi> by left environments E[la, a, lb]."""
    E = np.ones((1, 1, 1), dtype=np.complex128)
    for A, W in zip(sites, mpo):
        T = np.tensordot(E, A, axes=(2, 0))                      # (x, a, p, s)
        T = np.tensordot(T, W, axes=((1, 2), (0, 2)))            # (x, s, q, b)
        E = np.tensordot(np.conj(A), T, axes=((0, 1), (0, 2)))   # (r, s, b)
        E = np.transpose(E, (0, 2, 1))                           # (r, b, s)
    return complex(E[0, 0, 0])
