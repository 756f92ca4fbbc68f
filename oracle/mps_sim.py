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
        E = np.einsum("xy,xpa,ypb->ab", E, np.conj(A), B)
    return complex(E[0, 0])


def to_vector(sites):
    """Dense state, site 1 = fastest-varying bit (only for small N)."""
    psi = np.ones((1, 1), dtype=np.complex128)  # (phys..., bond)
    for A in sites:
        psi = np.tensordot(psi, A, axes=(psi.ndim - 1, 0))
    psi = psi.reshape(psi.shape[1:-1])
    return np.reshape(psi, (-1,), order="F")
