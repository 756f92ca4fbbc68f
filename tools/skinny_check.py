"""Parity of the tall-skinny GEMM tile variants (QTN_SKINNY=0/1/2, csrc/exec.cu:launch_gemm) against numpy.
Run once per setting: the variant is latched at the first launch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

q = graft.load_package()
rng = np.random.default_rng(1)
worst = 0.0
for M in (4096, 5000, 65536 + 3):
    for N in (1, 8, 9, 16, 17, 32, 33):
        for K in (1, 4, 16, 33, 300):
            a = rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))
            b = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
            got = q.ncon([a, b], [[-1, 1], [1, -2]])
            want = a @ b
            err = np.abs(got - want).max() / np.abs(want).max()
            worst = max(worst, err)
            assert err < 1e-12, (M, N, K, err)
            # gathered operand: contract over the FIRST (fastest) leg of a rank-3 tensor
            if K <= 33 and M <= 5000:
                t = rng.standard_normal((K, M // 8, 8)) + 1j * rng.standard_normal((K, M // 8, 8))
                got = q.ncon([t, b], [[1, -1, -2], [1, -3]])
                want = np.einsum("kij,kn->ijn", t, b)
                err = np.abs(got - want).max() / np.abs(want).max()
                assert err < 1e-12, ("gather", M, N, K, err)
# B stored (N, K): its contiguous direction is n, A's is k (both stage-load mappings in one launch), split-K shapes
for (M, N, K) in ((1024, 32, 1 << 16), (16, 16, 1 << 18), (4096, 24, 40), (300, 7, 5000)):
    a = rng.standard_normal((K, M)) + 1j * rng.standard_normal((K, M))
    b = rng.standard_normal((N, K)) + 1j * rng.standard_normal((N, K))
    got = q.ncon([a, b], [[1, -1], [-2, 1]])
    want = a.T @ b.T
    err = np.abs(got - want).max() / np.abs(want).max()
    worst = max(worst, err)
    assert err < 1e-12, ("kmajor", M, N, K, err)
print("QTN_SKINNY=%s ok, worst rel err %.2e, launches %d" % (os.environ.get("QTN_SKINNY", "1 (default)"), worst, q.launch_count()))
