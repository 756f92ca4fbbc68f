"""``contract`` and friends (src/contract.jl) on top of the CUDA plan executor.

The arithmetic the reference delegates to ``TensorOperations.ncon``
(src/contract.jl:257, 263) runs in ``libqaintensor_cuda`` (gather-GEMMs on the FP64
tensor pipe); this module only assigns labels and marshals buffers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NetworkArgs, arr_i32, arr_i64, as_c128, check, data_ptrs, lib


def contract_rep(net, optimize=False):
    """Label assignment of src/contract.jl:8-32 / :39-60: +k for contraction k,
    -(i + ncontractions) for open leg i."""
    shapes = [t.data.shape for t in net.tensors]     # one attribute walk per tensor (this runs on every contract(net))
    indexlist = [[0] * len(sh) for sh in shapes]
    leg_costs = {}
    for k, s in enumerate(net.contractions, 1):
        for (t, l) in s.idx:
            sh = shapes[t - 1]
            if not 1 <= l <= len(sh):
                raise AssertionError("leg %d out of range for tensor %d" % (l, t))
            indexlist[t - 1][l - 1] = k
            if optimize:
                leg_costs[k] = sh[l - 1]
    nc = len(net.contractions)
    for i, (t, l) in enumerate(net.openidx, 1):
        if indexlist[t - 1][l - 1] != 0:
            raise AssertionError("open leg participates in a contraction")
        indexlist[t - 1][l - 1] = -i - nc
        if optimize:
            leg_costs[i + nc] = shapes[t - 1][l - 1]
    for idx in indexlist:
        if 0 in idx:
            raise AssertionError("tensor leg without contraction or open index")
    return (leg_costs, indexlist) if optimize else indexlist


def contract_order(net, leg_costs, indexlist):
    """Exhaustive cost-capped order search (src/contract.jl:184-235), host C++."""
    nt = len(indexlist)
    nl = len(leg_costs)
    args = NetworkArgs([t.size() for t in net.tensors], indexlist)
    legdims = arr_i64([leg_costs[i] for i in range(1, nl + 1)])
    seq = (C.c_int32 * max(nl, 1))()
    nseq = C.c_int32(0)
    cost = C.c_int64(0)
    check(lib.qtn_order_exhaustive(nt, args.ranks, args.labels, nl, legdims, seq, C.byref(nseq), C.byref(cost)))
    return [int(seq[i]) for i in range(nseq.value)], int(cost.value)


class ContractionPlan:
    """Reusable plan: ``qtn_plan_create`` ... ``qtn_plan_destroy``."""

    def __init__(self, shapes, labels, order=None, slice_labels=(), precision="c128"):
        self.dtype = _lib.dtype_code(precision)
        self.np_dtype = np.complex64 if self.dtype == _lib.QTN_C64 else np.complex128
        self.shapes = [tuple(int(d) for d in s) for s in shapes]
        self.labels = [list(l) for l in labels]
        self._args = NetworkArgs(self.shapes, self.labels)
        self._h = C.c_void_p()
        ord_arr = arr_i32(order) if order is not None else None
        check(lib.qtn_plan_create(self._args.nt, self._args.ranks, self._args.dims, self._args.labels, ord_arr,
                                  len(order) if order is not None else 0, arr_i32(list(slice_labels)),
                                  len(slice_labels), self.dtype, C.byref(self._h)))
        info = (C.c_int64 * 8)()
        cost = (C.c_double * 2)()
        check(lib.qtn_plan_info(self._h, info, cost))
        self.nsteps, self.nslices, self.out_rank, self.out_numel = (int(info[i]) for i in range(4))
        self.max_elems, self.n_invariant, self.arena_bytes, self.launches_per_slice = (int(info[i]) for i in range(4, 8))
        self.flops_per_slice, self.bytes_per_slice = float(cost[0]), float(cost[1])
        od = (C.c_int64 * max(self.out_rank, 1))()
        check(lib.qtn_plan_out_dims(self._h, od))
        self.out_dims = tuple(int(od[i]) for i in range(self.out_rank))
        self._keep = None

    def steps(self):
        mnk = (C.c_int64 * (3 * max(self.nsteps, 1)))()
        flags = (C.c_int32 * max(self.nsteps, 1))()
        check(lib.qtn_plan_steps(self._h, mnk, flags))
        return [(int(mnk[3 * i]), int(mnk[3 * i + 1]), int(mnk[3 * i + 2]), int(flags[i])) for i in range(self.nsteps)]

    def n_pairwise(self):
        """Number of pairwise (GEMM) steps, i.e. without trace steps and operand pre-permutes."""
        return sum(1 for (_, _, _, f) in self.steps() if (f >> 1) & 7 == 0)

    def _marshal(self, arrays):
        """Arrays in the ABI's layout plus their pointer table; repeated calls with the very same
        (already column-major, right-precision) array objects re-use the marshalled table."""
        key = tuple(id(a) for a in arrays)
        cached = getattr(self, "_marshal_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        arrs = [_lib.as_cx(a, self.dtype) for a in arrays]
        for a, s in zip(arrs, self.shapes):
            if tuple(a.shape) != s:
                raise ValueError("tensor shape %r does not match the plan's %r" % (a.shape, s))
        if all(x is y for x, y in zip(arrs, arrays)):  # no conversion copies: pointers stay valid with the inputs
            self._marshal_cache = (key, arrs, data_ptrs(arrs))
        else:
            self._marshal_cache = None
        return arrs

    def _ptrs(self, arrs):
        cached = getattr(self, "_marshal_cache", None)
        return cached[2] if cached is not None and cached[1] is arrs else data_ptrs(arrs)

    def upload(self, arrays):
        _lib.require_device()
        arrs = self._marshal(arrays)
        check(lib.qtn_plan_upload(self._h, self._ptrs(arrs)))

    def execute_device(self, dev_ptr, slice_begin=0, slice_end=None):
        """Accumulate slices into a caller-owned device buffer (asynchronous)."""
        check(lib.qtn_plan_execute(self._h, slice_begin, self.nslices if slice_end is None else slice_end,
                                   C.c_void_p(dev_ptr)))

    def execute(self, arrays=None, slice_begin=0, slice_end=None):
        """Host in, host out (sum over the requested slices)."""
        _lib.require_device()
        out = np.zeros(self.out_dims, dtype=self.np_dtype, order="F")
        ptrs = None
        if arrays is not None:
            arrs = self._marshal(arrays)
            ptrs = self._ptrs(arrs)
        check(lib.qtn_plan_execute_host(self._h, ptrs, slice_begin, self.nslices if slice_end is None else slice_end,
                                        out.ctypes.data_as(C.c_void_p)))
        return out

    def contract_sliced(self, arrays, rank=0, nranks=1, first_slice=0, nslices=None):
        """Slice-parallel contraction of the window [first_slice, first_slice + nslices)
        (default: all slices): this rank's contiguous block + one NCCL allreduce."""
        _lib.require_device()
        out = np.zeros(self.out_dims, dtype=self.np_dtype, order="F")
        ptrs = None
        if arrays is not None:
            arrs = self._marshal(arrays)
            ptrs = self._ptrs(arrs)
        check(lib.qtn_contract_sliced_range(self._h, ptrs, first_slice, self.nslices if nslices is None else nslices,
                                            rank, nranks, out.ctypes.data_as(C.c_void_p)))
        return out

    def time_steps(self, slice_id=0):
        ms = (C.c_float * max(self.nsteps, 1))()
        check(lib.qtn_plan_time_steps(self._h, slice_id, ms))
        return [float(ms[i]) for i in range(self.nsteps)]

    def close(self):
        if self._h:
            lib.qtn_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def choose_slices(shapes, labels, order=None, max_log2_elems=28, min_slices=1, allow_partial=False):
    """Deterministic greedy slice-label choice (EXTENSION; rule in DESIGN.md).  Raises ``QtnError`` (QTN_EDOMAIN) when
    the target cannot be met -- the largest tensor has only open / extent-1 labels left, or fewer than ``min_slices``
    slices exist; ``allow_partial=True`` returns the labels found up to that point instead."""
    args = NetworkArgs(shapes, labels)
    ncap = max(sum(len(l) for l in labels), 1)
    out = (C.c_int32 * ncap)()
    n = C.c_int32(0)
    ord_arr = arr_i32(order) if order is not None else None
    rc = lib.qtn_choose_slices(args.nt, args.ranks, args.dims, args.labels, ord_arr,
                               len(order) if order is not None else 0, max_log2_elems, min_slices, out, C.byref(n))
    found = [int(out[i]) for i in range(n.value)]
    if rc == _lib.QTN_EDOMAIN and allow_partial:
        return found
    if rc == _lib.QTN_EDOMAIN:  # target not reachable (largest tensor has only open / extent-1 labels left, or too few slices)
        err = _lib.QtnError(rc, lib.qtn_last_error().decode("utf-8", "replace"))
        err.partial_labels = found   # what the rule found before it ran out of labels
        raise err
    check(rc)
    return found


def search_order(shapes, labels, ntrials=256, seed=0, max_log2_elems=-1):
    """Randomised greedy order search (EXTENSION; ``qtn_order_search``).  Returns
    ``(order, info)``: a complete label sequence for ``order=`` and the exact planner cost of it
    (total flops over all slices, flops per slice, slices, log2 of the largest tensor)."""
    args = NetworkArgs(shapes, labels)
    ncap = max(sum(len(l) for l in labels), 1)
    out = (C.c_int32 * ncap)()
    n = C.c_int32(0)
    cost = (C.c_double * 4)()
    check(lib.qtn_order_search(args.nt, args.ranks, args.dims, args.labels, int(ntrials), int(seed), int(max_log2_elems),
                               out, C.byref(n), cost))
    info = {"total_flops": float(cost[0]), "flops_per_slice": float(cost[1]), "nslices": float(cost[2]),
            "log2_max_elems": float(cost[3])}
    return [int(out[i]) for i in range(n.value)], info


def ncon(arrays, indexlist, order=None, precision="c128"):
    """``TensorOperations.ncon(tensors, indexlist; order)`` on the GPU (one-shot).
    ``precision="c64"`` selects the optional ComplexF32 mode (EXTENSION vii)."""
    _lib.require_device()
    code = _lib.dtype_code(precision)
    arrs = [_lib.as_cx(a, code) for a in arrays]
    args = NetworkArgs([a.shape for a in arrs], indexlist)
    nopen = sum(1 for l in indexlist for x in l if x < 0)
    out_n = 1
    for a, l in zip(arrs, indexlist):
        for d, x in zip(a.shape, l):
            if x < 0:
                out_n *= d
    out = np.zeros(max(out_n, 1), dtype=np.complex64 if code == _lib.QTN_C64 else np.complex128)
    rank = C.c_int32(0)
    dims = (C.c_int64 * 64)()
    if nopen > 64:
        raise ValueError("more than 64 open legs")
    ord_arr = arr_i32(order) if order is not None else None
    if len(arrs) > 32 and sum(a.size for a in arrs) <= (1 << 20):
        keep, ptrs = _lib.packed_ptrs(arrs)   # many small tensors (a circuit's gates)
    else:
        ptrs = data_ptrs(arrs)
    check(lib.qtn_contract(args.nt, ptrs, args.ranks, args.dims, args.labels, ord_arr,
                           len(order) if order is not None else 0, code, out.ctypes.data_as(C.c_void_p),
                           C.byref(rank), dims))
    shape = tuple(int(dims[i]) for i in range(rank.value))
    return np.reshape(out[:out_n], shape, order="F")


def permutedims(a, perm):
    """Julia ``permutedims(a, perm)`` (1-based perm) on the GPU."""
    _lib.require_device()
    a = as_c128(a)
    out_shape = tuple(a.shape[p - 1] for p in perm)
    out = np.zeros(out_shape, dtype=np.complex128, order="F")
    check(lib.qtn_permutedims(a.ctypes.data_as(C.c_void_p), a.ndim, arr_i64(a.shape), arr_i32(perm), _lib.QTN_C128,
                              out.ctypes.data_as(C.c_void_p)))
    return out


def contract(net, optimize=False, precision="c128", max_log2_elems=None, min_slices=1, rank=0, nranks=1):
    """``contract(net::TensorNetwork, optimize::Bool=false)`` (src/contract.jl:242-264).

    EXTENSION keywords with reference-preserving defaults: ``precision`` ("c128" | "c64");
    ``max_log2_elems`` slices the contraction until no tensor exceeds 2^k elements (deterministic greedy
    rule) and, with ``nranks > 1`` (after ``qtn_nccl_init``), gives every rank a contiguous block of the
    slices and sums the partial results with one NCCL allreduce."""
    if max_log2_elems is not None and len(net.tensors) > 1 and not optimize:
        il = contract_rep(net)
        arrays = [t.data for t in net.tensors]
        shapes = [a.shape for a in arrays]
        S = choose_slices(shapes, il, None, max_log2_elems, min_slices)
        plan = ContractionPlan(shapes, il, None, S, precision=precision)
        try:
            return plan.contract_sliced(arrays, rank, nranks)
        finally:
            plan.close()
    if len(net.tensors) == 1:
        out = permutedims(net.tensors[0].data, [l for (_, l) in net.openidx])
        return out.astype(np.complex64) if _lib.dtype_code(precision) == _lib.QTN_C64 else out
    arrays = [t.data for t in net.tensors]
    if optimize:
        leg_costs, indexlist = contract_rep(net, True)
        sequence, _ = contract_order(net, leg_costs, indexlist)
        # src/contract.jl:250-257: labels renamed to their position in `sequence`,
        # and `order=sequence` passed on top (quirk kept for bit-exact order parity).
        for lab in indexlist:
            for j, x in enumerate(lab):
                if x > 0:
                    lab[j] = sequence.index(x) + 1
        return ncon(arrays, indexlist, order=sequence, precision=precision)
    return ncon(arrays, contract_rep(net), precision=precision)
