"""Device-resident MPS evolution (EXTENSION composed from src/switch.jl:18-56: two-site
theta -> gate -> SVD -> truncate by src/svd.jl:29-33 plus a max-bond cap -> split).
Site tensors use the OpenMPS layout (lbond, 2, rbond) of src/mps.jl:99-110."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_c128, check, lib


class DeviceMPS:
    def __init__(self, sites, capacity):
        _lib.require_device()
        arrs = [as_c128(s) for s in sites]
        for a in arrs:
            if a.ndim != 3 or a.shape[1] != 2:
                raise ValueError("MPS site tensors must have shape (lbond, 2, rbond)")
        self.n = len(arrs)
        self.capacity = int(capacity)
        self._h = C.c_void_p()
        ptrs = (C.c_void_p * self.n)(*[a.ctypes.data for a in arrs])
        check(lib.qtn_mps_create(self.n, ptrs, _lib.arr_i64([a.shape[0] for a in arrs]),
                                 _lib.arr_i64([a.shape[2] for a in arrs]), self.capacity, C.byref(self._h)))

    @classmethod
    def product_state(cls, nsites, capacity, vectors=None):
        sites = []
        for i in range(nsites):
            v = np.array([1.0, 0.0], dtype=np.complex128) if vectors is None else np.asarray(vectors[i], np.complex128)
            sites.append(v.reshape(1, 2, 1))
        return cls(sites, capacity)

    def bonds(self):
        lb = (C.c_int64 * self.n)()
        rb = (C.c_int64 * self.n)()
        check(lib.qtn_mps_bonds(self._h, lb, rb))
        return [int(x) for x in lb], [int(x) for x in rb]

    def download(self):
        lb, rb = self.bonds()
        outs = [np.zeros((l, 2, r), dtype=np.complex128, order="F") for l, r in zip(lb, rb)]
        ptrs = (C.c_void_p * self.n)(*[o.ctypes.data for o in outs])
        check(lib.qtn_mps_download(self._h, ptrs))
        return outs

    def apply_layer(self, sites, gates, er=0.0, maxdim=0):
        """Gates (4x4, index = p_site + 2*p_{site+1}) on disjoint bonds; one batched SVD."""
        g = np.ascontiguousarray(np.stack([np.asfortranarray(np.asarray(x, np.complex128)).reshape(-1, order="F") for x in gates]))
        disc = (C.c_double * len(sites))()
        check(lib.qtn_mps_apply_layer(self._h, len(sites), _lib.arr_i32(sites), g.ctypes.data_as(C.c_void_p), float(er),
                                      int(maxdim), disc))
        return [float(d) for d in disc]

    def apply_gate2(self, site, gate, er=0.0, maxdim=0):
        return self.apply_layer([site], [gate], er, maxdim)[0]

    def overlap(self, other):
        out = (C.c_double * 2)()
        check(lib.qtn_mps_overlap(self._h, other._h, out))
        return complex(out[0], out[1])

    def close(self):
        if self._h:
            lib.qtn_mps_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def brickwork_layer_sites(nsites, half):
    """1-based left sites of the gates of half-layer 0 (bonds 1,3,5,...) or 1 (bonds 2,4,...)."""
    return list(range(1 + half, nsites, 2))


# ---- MPO x MPS (EXTENSION iii / iv): site-tensor MPO, apply + compress, expectation value ------------
def tfi_mpo(n, J=1.0, h=1.0):
    """Transverse-field Ising H = -J sum Z_i Z_{i+1} - h sum X_i as a D = 3 MPO of site tensors
    (bond_in, out, in, bond_out) (layout of src/mpo.jl:66); the reference's own MPO constructor
    needs the dense 2^n matrix (src/mpo.jl:27) and its site-tensor constructor errors (:16-19)."""
    W = np.zeros((3, 2, 2, 3), dtype=np.complex128)
    W[0, :, :, 0] = np.eye(2)
    W[1, :, :, 0] = np.diag([1.0, -1.0])
    W[2, :, :, 0] = -h * np.array([[0, 1], [1, 0]])
    W[2, :, :, 1] = -J * np.diag([1.0, -1.0])
    W[2, :, :, 2] = np.eye(2)
    return [np.asfortranarray(W[(2 if i == 0 else 0):(3 if i == 0 else 3), :, :, :(1 if i == n - 1 else 3)]) for i in range(n)]


def _mpo_args(mpo):
    arrs = [as_c128(w) for w in mpo]
    for w in arrs:
        if w.ndim != 4 or w.shape[1] != 2 or w.shape[2] != 2:
            raise ValueError("MPO site tensors must have shape (bond_in, 2, 2, bond_out)")
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    return arrs, ptrs, _lib.arr_i64([a.shape[0] for a in arrs]), _lib.arr_i64([a.shape[3] for a in arrs])


def _apply_mpo(self, mpo, er=0.0, maxdim=0):
    """|psi> <- compress(MPO |psi>) with cutoff `er` and bond cap `maxdim` (default: capacity)."""
    if len(mpo) != self.n:
        raise ValueError("MPO and MPS lengths differ")
    arrs, ptrs, dl, dr = _mpo_args(mpo)
    disc = (C.c_double * max(self.n - 1, 1))()
    check(lib.qtn_mps_apply_mpo(self._h, ptrs, dl, dr, float(er), int(maxdim), disc))
    return [float(disc[i]) for i in range(self.n - 1)]


def _expect_mpo(self, mpo):
    if len(mpo) != self.n:
        raise ValueError("MPO and MPS lengths differ")
    arrs, ptrs, dl, dr = _mpo_args(mpo)
    out = (C.c_double * 2)()
    check(lib.qtn_mps_expect_mpo(self._h, ptrs, dl, dr, out))
    return complex(out[0], out[1])


DeviceMPS.apply_mpo = _apply_mpo
DeviceMPS.expect_mpo = _expect_mpo
