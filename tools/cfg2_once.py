import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
net, _, _ = q.circuits.cfg2_network()
q.optimize_contraction_order(net)
il = q.contract_rep(net)
arrays = [t.data for t in net.tensors]
plan = q.ContractionPlan([a.shape for a in arrays], il)
print(complex(plan.execute(arrays)))   # first call: direct launches in step order (then graph capture)
import json
json.dump(plan.steps(), open("gpurun_out/cfg2_plan_steps.json", "w"))
