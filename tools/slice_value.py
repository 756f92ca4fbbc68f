"""Value of one slice of the cfg-3 plan (reference order, or --search), for A/B runs of planner / kernel options
selected by environment variables (QTN_PREPERMUTE, QTN_SKINNY, QTN_GEMM_TILE, ...), which are latched per process."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

q = graft.load_package()
level = int(sys.argv[1]) if len(sys.argv) > 1 else 31
sid = int(sys.argv[2]) if len(sys.argv) > 2 else 99
search = "--search" in sys.argv
net, _, _ = q.circuits.cfg3_network()
arrays = [t.data for t in net.tensors]
shapes = [a.shape for a in arrays]
order = None
if search:
    il = q.contract_rep(net)
    order, _ = q.search_order(shapes, il, 512, 0, level)
else:
    q.optimize_contraction_order(net)
    il = q.contract_rep(net)
S = q.choose_slices(shapes, il, order, level, 1)
plan = q.ContractionPlan(shapes, il, order, S)
sid %= plan.nslices
plan.execute(arrays, sid, sid + 1)
t0 = time.perf_counter()
v = complex(plan.execute(None, sid, sid + 1))
dt = time.perf_counter() - t0
npre = sum(1 for (_, _, _, f) in plan.steps() if (f >> 1) & 7 == 1)
print("slice %d of %d level %d search=%s prepermutes=%d value=%r seconds=%.4f" % (sid, plan.nslices, level, search, npre, v, dt))
