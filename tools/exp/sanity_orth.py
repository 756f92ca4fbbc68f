"""Small driver for compute-sanitizer: CholeskyQR2 gauge step, its fallbacks and the null-vector completion."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as graft
q = graft.load_package()
svdmod = sys.modules["qaintensor_b200.svd"]
rng = np.random.default_rng(0)
cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
for m, n in ((300, 200), (130, 130)):
    A = cr(m, n)
    Q, method = svdmod.orth_columns(A)
    print("orth", m, n, method, np.abs(Q.conj().T @ Q - np.eye(n)).max(), np.abs(Q @ (Q.conj().T @ A) - A).max())
D = cr(200, 30) @ cr(30, 140)
Q, method = svdmod.orth_columns(D)
print("orth deficient", method, np.abs(Q.conj().T @ Q - np.eye(140)).max(), np.abs(Q @ (Q.conj().T @ D) - D).max())
for A in (cr(40, 6) @ cr(6, 30), cr(30, 6) @ cr(6, 40), np.zeros((5, 4), complex), np.zeros((3, 7), complex), cr(70, 1) @ cr(1, 70)):
    U, S, Vh, k = q.svd_trunc(A)
    r = min(A.shape)
    print("svd", A.shape, "k", k, "|U^H U - I|", np.abs(U.conj().T @ U - np.eye(r)).max(), "|Vh Vh^H - I|", np.abs(Vh @ Vh.conj().T - np.eye(r)).max(),
          "rec", np.abs((U * S) @ Vh - A).max())
