"""Parity of plans with operand pre-permutes (csrc/plan.cpp: merge) against the CPU oracle.  Run with
QTN_PREPERMUTE=2 so that every scattered B operand is pre-permuted, also in small networks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import rel_err, to_oracle  # noqa: E402
from oracle import contract as oc  # noqa: E402
from test_host_planner import _random_general_network  # noqa: E402

q = graft.load_package()
rng = np.random.default_rng(31)
npermutes = 0


def check(net, order=None, max_log2=None, tol=1e-10):
    global npermutes
    want = oc.contract(to_oracle(net))
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    shapes = [a.shape for a in arrays]
    S = q.choose_slices(shapes, il, order, max_log2, 2) if max_log2 is not None else []
    for prec, t in (("c128", tol), ("c64", 1e-4)):
        plan = q.ContractionPlan(shapes, il, order, S, precision=prec)
        npermutes += sum(1 for (_, _, _, f) in plan.steps() if (f >> 1) & 7 == 1)
        got = plan.execute(arrays)
        assert got.shape == np.shape(want) and rel_err(got, want) < t, (prec, rel_err(got, want))
        plan.close()


for trial in range(30):
    check(_random_general_network(q, rng, int(rng.integers(3, 9)), int(rng.integers(2, 12)), int(rng.integers(0, 4))))
for (nq, depth, seed) in ((12, 10, 1), (16, 12, 11)):
    net, _, _ = q.circuits.cfg2_network(nq, depth, seed=seed)
    check(net)                      # default order: long chains of scattered operands
    n2 = net.copy()
    q.optimize_contraction_order(n2)
    check(n2)
    check(n2, max_log2=8)
    il = q.contract_rep(net)
    order, _ = q.search_order([t.data.shape for t in net.tensors], il, 32, 0, -1)
    check(net, order=order)
net, _, _ = q.circuits.cfg3_network(4, 4, 8, seed=21)
il = q.contract_rep(net)
order, _ = q.search_order([t.data.shape for t in net.tensors], il, 32, 0, 7)
check(net, order=order, max_log2=7)
print("QTN_PREPERMUTE=%s ok, %d permute steps executed" % (os.environ.get("QTN_PREPERMUTE", "1"), npermutes))
