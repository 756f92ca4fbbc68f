import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
import qaintensor_b200.mps as pm
ps = sys.modules["qaintensor_b200.svd"]
rng = np.random.default_rng(8)
N = 6
bs = [rng.standard_normal(2) + 1j * rng.standard_normal(2) for _ in range(N)]
psi = bs[0]
for b in bs[1:]:
    psi = np.kron(psi, b)
orig = ps.contract_svd
def spy(T1, T2, idx, er=0.0):
    try:
        return orig(T1, T2, idx, er)
    except Exception as e:
        print("FAIL", T1.data.shape, T2.data.shape, idx, e)
        np.save("gpurun_out/fail_T1.npy", T1.data); np.save("gpurun_out/fail_T2.npy", T2.data)
        a = T1.data; b = T2.data
        A = np.reshape(a, (-1, a.shape[-1]), order="F"); B = np.reshape(b, (b.shape[0], -1), order="F")
        for nm, M in (("T1p", A), ("T2p", B)):
            try:
                U, S, Vh, k = q.svd_trunc(M); print(nm, M.shape, "single ok", S[:4])
            except Exception as e2:
                print(nm, M.shape, "single FAIL", e2, np.linalg.svd(M, compute_uv=False)[:4])
        raise
pm.contract_svd = spy
for order in ([2, 1, 3, 4, 5, 6], [3, 1, 6, 2, 5, 4]):
    m = q.MPS(psi)
    q.permute(m, order)
    print("ok", order)
