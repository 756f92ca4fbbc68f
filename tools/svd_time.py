"""Times the batched Jacobi SVD through the C ABI (host buffers) and checks it against LAPACK:
    python tools/svd_time.py [batch] [m] [n] [reps]
QTN_JACOBI_STATS=1 prints the per-call sweep statistics and the device time of the sweeps (stderr);
QTN_JACOBI=fused selects the previous fused round kernel for A/B runs."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

q = graft.load_package()
from qaintensor_b200 import _lib  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 24
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n = int(sys.argv[3]) if len(sys.argv) > 3 else m
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
_lib.require_device()
rng = np.random.default_rng(5)
mats = [np.asfortranarray(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) for _ in range(batch)]
r = min(m, n)
Us = [np.zeros((m, r), complex, order="F") for _ in range(batch)]
Ss = [np.zeros(r) for _ in range(batch)]
Vs = [np.zeros((r, n), complex, order="F") for _ in range(batch)]
ks = (C.c_int64 * batch)()
vp = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])  # noqa: E731
for rep in range(reps):
    t0 = time.perf_counter()
    _lib.check(_lib.lib.qtn_svd_trunc_batched(batch, vp(mats), _lib.arr_i64([m] * batch), _lib.arr_i64([n] * batch), -1.0, 0,
                                              vp(Us), vp(Ss), vp(Vs), ks))
    print("rep %d: %.1f ms wall (incl. H2D/D2H)" % (rep, 1e3 * (time.perf_counter() - t0)))
worst = 0.0
for A, U, S, Vh in list(zip(mats, Us, Ss, Vs))[:min(batch, 3)]:
    Sref = np.linalg.svd(A, compute_uv=False)
    worst = max(worst, np.abs(S - Sref).max() / Sref[0], np.abs((U * S) @ Vh - A).max() / Sref[0],
                np.abs(U.conj().T @ U - np.eye(r)).max(), np.abs(Vh @ Vh.conj().T - np.eye(r)).max())
print("max deviation (sigma, reconstruction, orthogonality) over the checked matrices: %.2e" % worst)
assert worst < 1e-10
