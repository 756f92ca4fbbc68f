"""ctypes binding of ``libqaintensor_cuda.so`` (C ABI in ``include/qaintensor_cuda.h``).

There is no CPU fallback: importing works without a GPU (host-only entry points such
as ordering and planning are usable), but every compute entry point raises
``QtnError`` when the library or an sm_100 device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqaintensor_cuda.so")

QTN_C128 = 0
QTN_C64 = 1
QTN_ENODEVICE = -2
QTN_EDOMAIN = -6
QTN_EBUSY = -7


class QtnError(RuntimeError):
    """Raised for every non-zero return of the C ABI; ``.code`` holds the QTN_E* value."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise QtnError(QTN_ENODEVICE, "libqaintensor_cuda.so not built at %s -- run __graft_entry__.build(); "
                       "there is no CPU fallback" % LIB_PATH)
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER
lib.qtn_last_error.restype = C.c_char_p
lib.qtn_stream.restype = vp
lib.qtn_launch_count.restype = i64
lib.qtn_launch_count.argtypes = [C.c_int]

_SIGS = {
    "qtn_init": [C.c_int],
    "qtn_device_count": [P(C.c_int)],
    "qtn_bench_dmma_peak": [P(f64)],
    "qtn_order_treewidth": [i32, i32, P(i32), P(i32), P(i32)],
    "qtn_graph_treewidth": [i32, i32, P(i32), P(i32), P(i32)],
    "qtn_order_exhaustive": [i32, P(i32), P(P(i32)), i32, P(i64), P(i32), P(i32), P(i64)],
    "qtn_plan_create": [i32, P(i32), P(P(i64)), P(P(i32)), P(i32), i32, P(i32), i32, i32, P(vp)],
    "qtn_plan_destroy": [vp],
    "qtn_choose_slices": [i32, P(i32), P(P(i64)), P(P(i32)), P(i32), i32, i32, i64, P(i32), P(i32)],
    "qtn_order_search": [i32, P(i32), P(P(i64)), P(P(i32)), i32, C.c_uint64, i32, P(i32), P(i32), P(f64)],
    "qtn_plan_info": [vp, P(i64), P(f64)],
    "qtn_plan_out_dims": [vp, P(i64)],
    "qtn_plan_steps": [vp, P(i64), P(i32)],
    "qtn_plan_upload": [vp, P(vp)],
    "qtn_plan_execute": [vp, i64, i64, vp],
    "qtn_plan_execute_host": [vp, P(vp), i64, i64, vp],
    "qtn_plan_time_steps": [vp, i64, P(C.c_float)],
    "qtn_contract": [i32, P(vp), P(i32), P(P(i64)), P(P(i32)), P(i32), i32, i32, vp, P(i32), P(i64)],
    "qtn_nccl_unique_id": [vp],
    "qtn_nccl_init": [i32, i32, vp],
    "qtn_nccl_allreduce_sum_f64": [vp, i64],
    "qtn_nccl_allreduce_sum_f32": [vp, i64],
    "qtn_contract_sliced": [vp, P(vp), i32, i32, vp],
    "qtn_contract_sliced_range": [vp, P(vp), i64, i64, i32, i32, vp],
    "qtn_permutedims": [vp, i32, P(i64), P(i32), i32, vp],
    "qtn_permutedims_device": [vp, i32, P(i64), P(i32), i32, vp],
    "qtn_zgemm_device": [C.c_char, C.c_char, i64, i64, i64, vp, i64, vp, i64, vp, i64],
    "qtn_svd_trunc": [vp, i64, i64, f64, i64, vp, P(f64), vp, P(i64)],
    "qtn_svd_trunc_batched": [i32, P(vp), P(i64), P(i64), f64, i64, P(vp), P(vp), P(vp), P(i64)],
    "qtn_svd_trunc_device": [vp, i64, i64, f64, i64, vp, vp, vp, P(i64), P(i32)],
    "qtn_contract_svd": [vp, i32, P(i64), i32, vp, i32, P(i64), i32, f64, vp],
    "qtn_mps_create": [i32, P(vp), P(i64), P(i64), i64, P(vp)],
    "qtn_mps_destroy": [vp],
    "qtn_mps_bonds": [vp, P(i64), P(i64)],
    "qtn_mps_download": [vp, P(vp)],
    "qtn_mps_apply_gate2": [vp, i32, vp, f64, i64, P(f64)],
    "qtn_mps_apply_layer": [vp, i32, P(i32), vp, f64, i64, P(f64)],
    "qtn_mps_overlap": [vp, vp, P(f64)],
    "qtn_mps_from_vector": [vp, i32, P(vp), P(i64)],
    "qtn_mpo_from_matrix": [vp, i32, P(vp), P(i64)],
    "qtn_decompose": [vp, i32, P(vp), P(i64)],
    "qtn_contract_svd_fold": [i32, P(vp), P(i64), P(i64), P(i64), f64, vp, i64],
    "qtn_mps_switch_adjacent": [vp, i64, i64, vp, i64, vp, vp, P(i64)],
    "qtn_mps_apply_mpo": [vp, P(vp), P(i64), P(i64), f64, i64, P(f64)],
    "qtn_mps_expect_mpo": [vp, P(vp), P(i64), P(i64), P(f64)],
    "qtn_orth_columns": [vp, i64, i64, vp, P(i32)],
    "qtn_pool_stats": [P(i64)],
    "qtn_net_create": [i32, P(vp), P(i32), P(P(i64)), i32, P(i32), i32, P(i32), P(vp)],
    "qtn_net_destroy": [vp],
    "qtn_net_sizes": [vp, P(i32)],
    "qtn_net_structure": [vp, P(i32), P(i32)],
    "qtn_net_tensor": [vp, i32, P(i32), P(i64), P(vp)],
    "qtn_net_tensor_circuit": [vp, i32, P(i32), P(i32), P(vp)],
    "qtn_net_apply_mpo": [vp, vp, i32, P(i32), P(vp)],
    "qtn_net_extend_mpo": [vp, i32, P(i32)],
    "qtn_net_close": [vp, P(i32)],
    "qtn_net_optimize_order": [vp, i32, i32, C.c_uint64, i32],
    "qtn_net_contract": [vp, i32, i32, vp, P(i32), P(i64)],
}
for _name, _args in _SIGS.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = C.c_int


def check(rc):
    if rc != 0:
        raise QtnError(rc, lib.qtn_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    lib.qtn_device_count(C.byref(n))
    return n.value


_device_ok = False


def require_device():
    """Fail loudly when the CUDA path cannot run (no fallback exists)."""
    global _device_ok
    if _device_ok:
        return
    check(lib.qtn_init(int(os.environ.get("LOCAL_RANK", "0")) if device_count() > 1 else 0))
    _device_ok = True


def dmma_peak_tflops():
    """Measured FP64 tensor-pipe ceiling (register-only DMMA loop), TFLOP/s."""
    v = f64(0.0)
    check(lib.qtn_bench_dmma_peak(C.byref(v)))
    return float(v.value)


def pool_stats():
    """Workspace pool counters: (cudaMalloc calls, cache hits, bytes owned, blocks handed out)."""
    out = (i64 * 4)()
    check(lib.qtn_pool_stats(out))
    return tuple(int(x) for x in out)


def stream_ptr():
    return int(lib.qtn_stream() or 0)


def launch_count(reset=False):
    return int(lib.qtn_launch_count(1 if reset else 0))


def as_c128(a):
    """Column-major ComplexF64 copy/view of ``a`` (the ABI's memory layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.complex128))


def as_cx(a, dtype_code):
    """Column-major complex array in the precision of `dtype_code` (QTN_C128 / QTN_C64)."""
    return np.asfortranarray(np.asarray(a, dtype=np.complex64 if dtype_code == QTN_C64 else np.complex128))


def dtype_code(precision):
    if precision in (None, "c128", "ComplexF64", QTN_C128):
        return QTN_C128
    if precision in ("c64", "ComplexF32", QTN_C64):
        return QTN_C64
    raise ValueError("precision must be 'c128' (ComplexF64) or 'c64' (ComplexF32 mode)")


def arr_i32(v):
    return (i32 * max(len(v), 1))(*[int(x) for x in v])


def arr_i64(v):
    return (i64 * max(len(v), 1))(*[int(x) for x in v])


class NetworkArgs:
    """Marshals (tensors' ranks, dims, labels) into the ABI's pointer arrays: two flat buffers (dims, labels) and two
    pointer tables computed from their base addresses -- no per-tensor ctypes objects (a 260-tensor network costs
    0.2 ms instead of 0.9 ms per `contract(net)`)."""

    def __init__(self, shapes, labels):
        nt = len(shapes)
        self.nt = nt
        ranks = np.fromiter((len(s) for s in shapes), dtype=np.int32, count=nt)
        if any(len(l) != r for l, r in zip(labels, ranks)):
            raise ValueError("every tensor needs one label per leg")
        tot = int(ranks.sum())
        self._ranks = np.ascontiguousarray(ranks if nt else np.zeros(1, np.int32))
        self._dims = np.fromiter((d for s in shapes for d in s), dtype=np.int64, count=tot) if tot else np.zeros(1, np.int64)
        self._labs = np.fromiter((x for l in labels for x in l), dtype=np.int32, count=tot) if tot else np.zeros(1, np.int32)
        start = np.zeros(max(nt, 1), dtype=np.int64)
        if nt > 1:
            np.cumsum(ranks[:-1], out=start[1:nt])
        self._dptr = (self._dims.ctypes.data + 8 * start).astype(np.uint64)
        self._lptr = (self._labs.ctypes.data + 4 * start).astype(np.uint64)
        self.ranks = self._ranks.ctypes.data_as(P(i32))
        self.dims = self._dptr.ctypes.data_as(P(P(i64)))
        self.labels = self._lptr.ctypes.data_as(P(P(i32)))


def packed_ptrs(arrays):
    """Pointer table for many small column-major arrays of one dtype: they are copied into ONE packed buffer and the
    table is computed from its base address (0.6 -> 0.2 ms for the 260 tensors of a 20-qubit circuit, against one
    ctypes pointer object per tensor).  Memory order ("K") of a column-major array is the ABI's element order.
    Returns (objects to keep alive during the call, POINTER(c_void_p))."""
    flat = np.concatenate([a.ravel(order="K") for a in arrays])
    offs = np.zeros(len(arrays), dtype=np.int64)
    if len(arrays) > 1:
        np.cumsum([a.size for a in arrays[:-1]], out=offs[1:])
    table = (flat.ctypes.data + flat.itemsize * offs).astype(np.uint64)
    return (flat, table), table.ctypes.data_as(P(vp))


def data_ptrs(arrays):
    return (vp * max(len(arrays), 1))(*[a.__array_interface__["data"][0] for a in arrays])
