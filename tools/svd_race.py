import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
rng = np.random.default_rng(1)
for shape in ((40, 30), (100, 100), (48, 200)):
    A = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    U, S, Vh, k = q.svd_trunc(A, 1e-9)
    print(shape, "rec err", np.abs((U * S) @ Vh - A).max())
