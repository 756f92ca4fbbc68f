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
