"""SVD timing on the GPU box: single and batched 2chi x 2chi problems (diagnostics)."""
import os, sys, time
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
from qaintensor_b200 import _lib
import scipy.linalg
rng = np.random.default_rng(0)
cases = ((256, 1), (512, 1), (1024, 1), (512, 24), (1024, 24)) if len(sys.argv) < 3 else ((int(sys.argv[1]), int(sys.argv[2])),)
for n, batch in cases:
    mats = [np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) for _ in range(batch)]
    Us = [np.zeros((n, n), complex, order="F") for _ in range(batch)]
    Ss = [np.zeros(n) for _ in range(batch)]
    Vs = [np.zeros((n, n), complex, order="F") for _ in range(batch)]
    ks = (C.c_int64 * batch)()
    vp = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    args = (batch, vp(mats), _lib.arr_i64([n] * batch), _lib.arr_i64([n] * batch), 1e-10, n // 2, vp(Us), vp(Ss), vp(Vs), ks)
    _lib.require_device()
    _lib.check(_lib.lib.qtn_svd_trunc_batched(*args))
    _lib.launch_count(True)
    t0 = time.perf_counter()
    _lib.check(_lib.lib.qtn_svd_trunc_batched(*args))
    dt = time.perf_counter() - t0
    nl = _lib.launch_count(True)
    t1 = time.perf_counter()
    Sref = scipy.linalg.svd(mats[0], full_matrices=False, lapack_driver="gesdd")[1]
    dc = time.perf_counter() - t1
    flops = batch * 4 * (14 * n ** 3 + 8 * n ** 3)
    print("n=%d batch=%d gpu %.1f ms (%.2f ms/SVD, incl H2D/D2H) launches %d  model %.2f TF/s | cpu gesdd %.1f ms/SVD | S err %.2e" % (
        n, batch, dt * 1e3, dt * 1e3 / batch, nl, flops / dt / 1e12, dc * 1e3, np.abs(Ss[0] - Sref).max() / Sref[0]))
