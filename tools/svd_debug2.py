import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
rng = np.random.default_rng(1)
A = rng.standard_normal((40, 6)) @ rng.standard_normal((6, 30)) + 0j
U, S, Vh, k = q.svd_trunc(A, 1e-9)
np.set_printoptions(linewidth=200, precision=3)
print("S", S)
print("V unit err", np.abs(Vh @ Vh.conj().T - np.eye(30)).max())
AV = A @ Vh.conj().T
print("A V - U S", np.abs(AV - U * S).max(), "per col", np.abs(AV - U * S).max(axis=0))
print("rec", np.abs((U * S) @ Vh - A).max())
print("U col norms", np.linalg.norm(U, axis=0))
