"""Timing of the cfg-5 building blocks on the GPU box: gauge step (CholeskyQR2 vs U-only Jacobi) on a 3072 x 1536
site matrix and the truncating SVD of a 1536 x 1024 one (host buffers, so H2D/D2H are included)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
svdmod = sys.modules["qaintensor_b200.svd"]
from qaintensor_b200 import _lib
rng = np.random.default_rng(0)
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3072, 1536)
A = np.asfortranarray(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))
only = sys.argv[3] if len(sys.argv) > 3 else ""
for mode in (("cholqr2",) if only == "cholonly" else ("cholqr2", "jacobi")):
    if mode == "jacobi":
        os.environ["QTN_ORTH"] = "jacobi"
    svdmod.orth_columns(A)
    _lib.launch_count(True)
    t0 = time.perf_counter()
    Q, method = svdmod.orth_columns(A)
    dt = time.perf_counter() - t0
    nl = _lib.launch_count(True)
    print("orth %dx%d %-8s method=%d  %.1f ms  launches %d  |Q^H Q - I| %.2e  |Q Q^H A - A| %.2e" % (
        m, n, mode, method, dt * 1e3, nl, np.abs(Q.conj().T @ Q - np.eye(n)).max(), np.abs(Q @ (Q.conj().T @ A) - A).max()))
os.environ.pop("QTN_ORTH", None)
if only == "cholonly":
    sys.exit(0)
B = np.asfortranarray(rng.standard_normal((n, 2 * n // 3)) + 1j * rng.standard_normal((n, 2 * n // 3)))
q.svd_trunc(B, 1e-10, n // 3)
t0 = time.perf_counter()
q.svd_trunc(B, 1e-10, n // 3)
print("svd_trunc %dx%d: %.1f ms" % (B.shape[0], B.shape[1], (time.perf_counter() - t0) * 1e3))
