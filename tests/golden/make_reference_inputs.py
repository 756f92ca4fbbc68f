"""Writes tests/golden/reference_inputs_r02.txt: the seeded networks / matrices that `julia/gen_golden.jl` feeds to
the UNMODIFIED reference (Qaintensor.jl) on a machine that has Julia, to produce tests/golden/reference_r02.json.
Line-oriented so that Julia needs no JSON package to read it:

    network <name> <ntensors> <ncontractions> <with_data>
    tensor <rank> <d1> .. <dr> [re im re im ...]          (column-major; data only when with_data = 1)
    contraction <t1> <l1> <t2> <l2>                        (1-based tensor => leg pairs, src/tensor_network.jl:7-10)
    matrix <name> <m> <n> re im ...                        (column-major)

Run from the repo root: python tests/golden/make_reference_inputs.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import circuits as ocirc, network as on  # noqa: E402


def fmt(a):
    a = np.asarray(a, dtype=np.complex128).reshape(-1, order="F")
    return " ".join("%.17g %.17g" % (z.real, z.imag) for z in a)


def emit_network(f, name, net, with_data):
    f.write("network %s %d %d %d\n" % (name, len(net.tensors), len(net.contractions), int(with_data)))
    for t in net.tensors:
        shp = np.asarray(t.data).shape
        f.write("tensor %d %s%s\n" % (len(shp), " ".join(str(d) for d in shp), (" " + fmt(t.data)) if with_data else ""))
    for c in net.contractions:
        (t1, l1), (t2, l2) = c.idx
        f.write("contraction %d %d %d %d\n" % (t1, l1, t2, l2))


def random_network(rng, nt, ne, nself, npar):
    """random_TN of test/test_treewidth.jl:21-33 plus `nself` self-contractions and `npar` duplicated (parallel) edges."""
    nlegs = [0] * nt
    cons = []

    def leg(t):
        nlegs[t - 1] += 1
        return (t, nlegs[t - 1])
    for _ in range(ne):
        n1 = int(rng.integers(1, nt))
        n2 = int(rng.integers(n1 + 1, nt + 1))
        cons.append([leg(n1), leg(n2)])
    for _ in range(npar):
        (a, _), (b, _) = cons[int(rng.integers(0, len(cons)))]
        cons.insert(int(rng.integers(0, len(cons) + 1)), [leg(a), leg(b)])
    for _ in range(nself):
        t = int(rng.integers(1, nt + 1))
        cons.insert(int(rng.integers(0, len(cons) + 1)), [leg(t), leg(t)])
    tensors = [on.Tensor(rng.standard_normal((2,) * n) + 1j * rng.standard_normal((2,) * n)) if n else
               on.Tensor(rng.standard_normal((1,)) + 0j) for n in nlegs]
    return on.Network(tensors, [on.Summation(c) for c in cons], [])


def main():
    out = os.path.join(ROOT, "tests", "golden", "reference_inputs_r02.txt")
    with open(out, "w") as f:
        emit_network(f, "cfg2", ocirc.cfg2_network()[0], False)          # order only (484 contractions)
        emit_network(f, "cfg3", ocirc.cfg3_network()[0], False)          # order only (516 contractions)
        for seed in (1, 2, 3):                                          # the seeded 12-qubit circuits of golden_r01.json
            emit_network(f, "cfg2_12q_d10_seed%d" % seed, ocirc.cfg2_network(12, 10, seed=seed)[0], True)
        emit_network(f, "cfg3_4x4_c8_seed21", ocirc.cfg3_network(4, 4, 8, seed=21)[0], True)
        rng = np.random.default_rng(20261017 + 7)
        for i, (nt, ne, nself, npar) in enumerate([(6, 9, 1, 0), (6, 9, 0, 2), (8, 12, 2, 2), (10, 16, 1, 3), (12, 20, 2, 1), (9, 14, 0, 0)]):
            emit_network(f, "random_%d_self%d_par%d" % (i, nself, npar), random_network(rng, nt, ne, nself, npar), True)
        # test/test_svd.jl:53-83 spectrum (sigma_j = e^-j), 40 x 40 to keep the file small
        s = np.exp(-np.arange(40.0))
        for name in ("svd_exp_T1", "svd_exp_T2"):
            q1 = np.linalg.qr(rng.standard_normal((40, 40)) + 1j * rng.standard_normal((40, 40)))[0]
            q2 = np.linalg.qr(rng.standard_normal((40, 40)) + 1j * rng.standard_normal((40, 40)))[0]
            f.write("matrix %s 40 40 %s\n" % (name, fmt(q1 @ np.diag(s) @ q2)))
    print("written", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
