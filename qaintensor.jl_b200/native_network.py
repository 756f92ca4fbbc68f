"""``NativeNetwork``: thin handle on the library's own network builder (``qtn_net_*``, csrc/network.cpp).

The Python mirror (tensor_network.py, tensor_circuit.py, mpo.py) restates the reference's symbolic code for
the parity tests; this class drives the C++ restatement of the same functions, which is what a harness
without Python (C, C++, Julia ``ccall``) uses.  ``tests/test_native_network.py`` checks the two against
each other structurally (bit-exact) and against the oracle numerically."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NetworkArgs, arr_i32, as_c128, check, data_ptrs, lib
from .tensor_network import GeneralTensorNetwork, Summation, Tensor


class NativeNetwork:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_network(cls, net):
        """``GeneralTensorNetwork(tensors, contractions, openidx)`` (src/tensor_network.jl:26-33)."""
        arrs = [as_c128(t.data) for t in net.tensors]
        args = NetworkArgs([a.shape for a in arrs], [[0] * a.ndim for a in arrs])
        pairs = [x for s in net.contractions for p in s.idx for x in p]
        if any(len(s.idx) != 2 for s in net.contractions):
            raise ValueError("Contractions of more than 2 tensors not supported")
        opn = [x for p in net.openidx for x in p]
        h = C.c_void_p()
        check(lib.qtn_net_create(len(arrs), data_ptrs(arrs), args.ranks, args.dims, len(net.contractions), arr_i32(pairs),
                                 len(net.openidx), arr_i32(opn), C.byref(h)))
        return cls(h)

    def sizes(self):
        s = (C.c_int32 * 3)()
        check(lib.qtn_net_sizes(self._h, s))
        return int(s[0]), int(s[1]), int(s[2])

    def to_network(self):
        """Read the network back as the Python mirror's ``GeneralTensorNetwork``."""
        nt, nc, no = self.sizes()
        pairs = (C.c_int32 * max(4 * nc, 1))()
        opn = (C.c_int32 * max(2 * no, 1))()
        check(lib.qtn_net_structure(self._h, pairs, opn))
        tensors = []
        for i in range(1, nt + 1):
            rank = C.c_int32(0)
            dims = (C.c_int64 * 64)()
            ptr = C.c_void_p()
            check(lib.qtn_net_tensor(self._h, i, C.byref(rank), dims, C.byref(ptr)))
            shape = tuple(int(dims[j]) for j in range(rank.value))
            n = int(np.prod(shape)) if shape else 1
            buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 * n,)).copy()
            tensors.append(Tensor(np.reshape(buf.view(np.complex128), shape, order="F")))
        cons = [Summation([(int(pairs[4 * k]), int(pairs[4 * k + 1])), (int(pairs[4 * k + 2]), int(pairs[4 * k + 3]))])
                for k in range(nc)]
        return GeneralTensorNetwork(tensors, cons, [(int(opn[2 * i]), int(opn[2 * i + 1])) for i in range(no)])

    def tensor_circuit(self, gates):
        """``tensor_circuit!(psi, cgc)`` without decomposition (src/tensor_circuit.jl:44-51)."""
        mats = [as_c128(np.asarray(g.matrix)) for g in gates]
        nw = [len(g.iwire) for g in gates]
        wires = [w for g in gates for w in g.iwire]
        check(lib.qtn_net_tensor_circuit(self._h, len(gates), arr_i32(nw), arr_i32(wires), data_ptrs(mats)))

    def apply_mpo(self, op, iwire):
        """``apply_MPO(psi, mpo, iwire)`` (src/mpo.jl:232-252); returns a new network."""
        h = C.c_void_p()
        check(lib.qtn_net_apply_mpo(self._h, op._h, len(iwire), arr_i32(list(iwire)), C.byref(h)))
        return NativeNetwork(h)

    def extend_mpo(self, iwire):
        """``extend_MPO(mpo, iwire)`` (src/mpo.jl:122-157), in place like the reference."""
        check(lib.qtn_net_extend_mpo(self._h, len(iwire), arr_i32(list(iwire))))

    def close_wires(self, bits):
        check(lib.qtn_net_close(self._h, arr_i32([int(b) for b in bits])))

    def optimize_contraction_order(self, method="treewidth", ntrials=256, seed=0, max_log2_elems=-1):
        code = {"treewidth": 0, "search": 1}.get(method)
        if code is None:
            raise ValueError("method must be 'treewidth' (reference) or 'search' (extension)")
        check(lib.qtn_net_optimize_order(self._h, code, int(ntrials), int(seed), int(max_log2_elems)))

    def contract(self, precision="c128", max_log2_elems=-1):
        """``contract(net)`` (src/contract.jl:242-264) on the GPU."""
        _lib.require_device()
        code = _lib.dtype_code(precision)
        net = None
        nt, nc, no = self.sizes()
        # output size: product of the open legs' extents
        net = self.to_network() if no else None
        shape = tuple(net.tensors[t - 1].size()[l - 1] for (t, l) in net.openidx) if no else ()
        n = int(np.prod(shape)) if shape else 1
        out = np.zeros(n, dtype=np.complex64 if code == _lib.QTN_C64 else np.complex128)
        rank = C.c_int32(0)
        dims = (C.c_int64 * 64)()
        check(lib.qtn_net_contract(self._h, code, int(max_log2_elems), out.ctypes.data_as(C.c_void_p), C.byref(rank), dims))
        return np.reshape(out, tuple(int(dims[i]) for i in range(rank.value)), order="F")

    def destroy(self):
        if self._h:
            lib.qtn_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
