"""Where the one-shot `contract(net)` call spends its time (host wall-clock per phase, through the C ABI):
network marshalling, qtn_plan_create, qtn_plan_upload (device init + H2D), first / second qtn_plan_execute_host,
qtn_plan_destroy, and the all-in-one qtn_contract.  Networks: the reference's notebook benchmark (QFT-20)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft

q = graft.load_package()
from qaintensor_b200 import _lib
from qaintensor_b200._lib import NetworkArgs, arr_i32, data_ptrs, lib

_lib.require_device()
net = q.circuits.notebook_expectation_network(20)
for order in ("default", "optimized"):
    if order == "optimized":
        q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    arrays = [_lib.as_cx(t.data, _lib.QTN_C128) for t in net.tensors]
    shapes = [a.shape for a in arrays]
    for rep in range(3):
        t = [time.perf_counter()]
        args = NetworkArgs(shapes, il); t.append(time.perf_counter())
        h = C.c_void_p()
        _lib.check(lib.qtn_plan_create(args.nt, args.ranks, args.dims, args.labels, None, 0, arr_i32([]), 0, _lib.QTN_C128, C.byref(h))); t.append(time.perf_counter())
        ptrs = data_ptrs(arrays); t.append(time.perf_counter())
        _lib.check(lib.qtn_plan_upload(h, ptrs)); t.append(time.perf_counter())
        out = np.zeros((), dtype=np.complex128)
        _lib.check(lib.qtn_plan_execute_host(h, None, 0, 1, out.ctypes.data_as(C.c_void_p))); t.append(time.perf_counter())
        _lib.check(lib.qtn_plan_execute_host(h, None, 0, 1, out.ctypes.data_as(C.c_void_p))); t.append(time.perf_counter())
        _lib.check(lib.qtn_plan_execute_host(h, None, 0, 1, out.ctypes.data_as(C.c_void_p))); t.append(time.perf_counter())
        lib.qtn_plan_destroy(h); t.append(time.perf_counter())
        r = q.contract(net); t.append(time.perf_counter())
        names = ["NetworkArgs", "plan_create", "data_ptrs", "upload(+device init)", "execute_host #1", "execute_host #2", "execute_host #3", "destroy", "contract(net) one-shot"]
        print(order, "rep", rep, "  ".join("%s %.2f" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])), "ms")
