"""One contraction of the reference's notebook benchmark network (QFT-20 expectation value) per order, for launch
lists (`ncu --metrics gpu__time_duration.sum`): first executes are direct launches in step order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
net = q.circuits.notebook_expectation_network(20)
for order in ("default", "optimized"):
    if order == "optimized":
        q.optimize_contraction_order(net)
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    plan = q.ContractionPlan([a.shape for a in arrays], il)
    print(order, complex(plan.execute(arrays).reshape(-1)[0]), "steps", plan.nsteps, "launches/slice", plan.launches_per_slice)
    plan.close()
