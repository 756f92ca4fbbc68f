import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def q():
    """The product package (builds libqaintensor_cuda.so in-tree if stale)."""
    return graft.build()


@pytest.fixture(scope="session")
def gpu(q):
    """Fails (does not skip) when the CUDA path cannot run: there is no fallback to test."""
    from qaintensor_b200 import _lib
    _lib.require_device()
    return q


def to_oracle(net):
    from oracle import network as on
    return on.Network([on.Tensor(t.data) for t in net.tensors], [on.Summation(s.idx) for s in net.contractions],
                      list(net.openidx))


def random_TN(q, Nn, Ne, rng, complex_data=True):
    """`random_TN(Nn, Ne)` of test/test_treewidth.jl:21-33 with a seeded generator."""
    nlegs = [0] * Nn
    cons = []
    for _ in range(Ne):
        n1 = int(rng.integers(1, Nn))
        n2 = int(rng.integers(n1 + 1, Nn + 1))
        nlegs[n1 - 1] += 1
        nlegs[n2 - 1] += 1
        cons.append(q.Summation([(n1, nlegs[n1 - 1]), (n2, nlegs[n2 - 1])]))
    def rnd(shape):
        a = rng.standard_normal(shape)
        return a + 1j * rng.standard_normal(shape) if complex_data else a + 0j
    ts = [q.Tensor(rnd((2,) * n)) if n > 0 else q.Tensor(rnd((1,))) for n in nlegs]
    return q.GeneralTensorNetwork(ts, cons, [])


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
