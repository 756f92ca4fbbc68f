"""Parity of the persistent streaming kernel (csrc/kernels.cuh: zgemm_stream_kernel; M >= 16384, N <= 32, K <= 32)
against numpy: ragged M, every N / K class, both stage-load mappings (row-major and k-major A), gathered operands,
conjugated dense operands.  QTN_STREAM=0 runs the same shapes through the tile kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def run():
    q = graft.load_package()
    rng = np.random.default_rng(7)
    worst = 0.0

    def chk(got, want, what):
        nonlocal worst
        err = np.abs(got - want).max() / np.abs(want).max()
        worst = max(worst, err)
        assert err < 1e-12, (what, err)

    for M in (16384, 20000 + 7, 65536 + 3):
        for N in (1, 5, 8, 9, 16, 17, 32):
            for K in (1, 3, 4, 8, 13, 16, 20, 32):
                a = np.asfortranarray(rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K)))
                b = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
                chk(q.ncon([a, b], [[-1, 1], [1, -2]]), a @ b, ("rowmajor", M, N, K))
                if M <= 20007:
                    ak = np.asfortranarray(a.T)   # (K, M) column-major: k is the contiguous direction
                    chk(q.ncon([ak, b], [[1, -1], [1, -2]]), a @ b, ("kmajor", M, N, K))
    # dense intermediates (c_dense, plain store): the predicate-free epilogues, incl. the N = 4 half-block of a two-qubit gate
    for M in (16384, 32768 + 5):
        for N in (2, 3, 4, 8, 16):
            for K in (2, 4, 8):
                a = np.asfortranarray(rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K)))
                b = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
                c = rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3))
                chk(q.ncon([a, b, c], [[-1, 1], [1, 2], [2, -2]]), a @ b @ c, ("intermediate", M, N, K))
    # gathered operand: the contracted legs sit in the middle of a rank-5 tensor (runs of 4 contiguous rows)
    for (k1, k2, N) in ((2, 2, 4), (4, 4, 16), (2, 8, 8), (4, 8, 32)):
        t = np.asfortranarray(rng.standard_normal((4, k1, 64, k2, 128)) + 1j * rng.standard_normal((4, k1, 64, k2, 128)))
        b = rng.standard_normal((k1, k2, N)) + 1j * rng.standard_normal((k1, k2, N))
        chk(q.ncon([t, b], [[-1, 1, -2, 2, -3], [1, 2, -4]]), np.einsum("akbld,kln->abdn", t, b), ("gather", k1, k2, N))
    print("QTN_STREAM=%s ok, worst rel err %.2e, launches %d" % (os.environ.get("QTN_STREAM", "1 (default)"), worst, q.launch_count()))
    return worst


if __name__ == "__main__":
    run()
