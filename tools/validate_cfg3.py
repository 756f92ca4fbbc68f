"""Full-size parity chain for BASELINE config 3 (run on the GPU box; ~3 min).

  oracle (numpy, CPU):  one slice at slicing level 2^28  ==  sum of its 64 sub-slices at level 2^24
  GPU level 2^28 slice  vs  that oracle value                          (tolerance 1e-10 relative)
  GPU level 2^31 slice  ==  sum of its 32 GPU sub-slices at level 2^28 (greedy slice sets are nested)

so the large-tile GEMM shapes of the default bench (level 2^31) are tied back to the CPU oracle."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
q = graft.load_package()
from oracle import circuits as ocirc, contract as oc, network2graph as o2g, plan as op

net, _, _ = q.circuits.cfg3_network()
q.optimize_contraction_order(net)
il = q.contract_rep(net)
arrays = [t.data for t in net.tensors]
shapes = [a.shape for a in arrays]
S = {lvl: q.choose_slices(shapes, il, None, lvl, 1) for lvl in (24, 28, 31)}
assert S[28][:len(S[31])] == S[31] and S[24][:len(S[28])] == S[28], "greedy slice sets must be nested"
res = {"labels": {str(k): len(v) for k, v in S.items()}}

# --- oracle: level-28 slice `sid28` as the sum of its level-24 sub-slices -------------------------------
onet, _, _ = ocirc.cfg3_network()
o2g.optimize_contraction_order(onet)
oil = oc.contract_rep(onet)
oarr = [t.data for t in onet.tensors]
nodes, steps = op.contraction_tree(oil)
dims = op.label_dims(oarr, oil)
assert op.choose_slice_labels(nodes, steps, dims, 24, 1) == S[24]
sid28 = 12345
n28, n24 = 2 ** len(S[28]), 2 ** len(S[24])
t0 = time.time()
want = 0j
for t in range(n24 // n28):
    want += complex(op.execute_tree(oarr, oil, nodes, steps, op.slice_assignment(S[24], dims, sid28 + n28 * t)))
res["oracle_seconds"] = time.time() - t0
p28 = q.ContractionPlan(shapes, il, None, S[28])
got28 = complex(p28.execute(arrays, sid28, sid28 + 1))
res["slice28"] = {"sid": sid28, "oracle": [want.real, want.imag], "gpu": [got28.real, got28.imag], "rel_err": abs(got28 - want) / abs(want)}
print("level-28 slice vs oracle: rel err %.3e (oracle %.1f s)" % (res["slice28"]["rel_err"], res["oracle_seconds"]))

# --- GPU: level-31 slice == sum of its 32 level-28 sub-slices -------------------------------------------
sid31 = 777
n31 = 2 ** len(S[31])
sub = 0j
for t in range(n28 // n31):
    sub += complex(p28.execute(None, sid31 + n31 * t, sid31 + n31 * t + 1))
p28.close()
p31 = q.ContractionPlan(shapes, il, None, S[31])
got31 = complex(p31.execute(arrays, sid31, sid31 + 1))
res["slice31"] = {"sid": sid31, "sum_of_level28": [sub.real, sub.imag], "gpu": [got31.real, got31.imag], "rel_err": abs(got31 - sub) / abs(sub),
                  "arena_GB": p31.arena_bytes / 1e9}
print("level-31 slice vs sum of 32 level-28 slices: rel err %.3e" % res["slice31"]["rel_err"])
json.dump(res, open("gpurun_out/validate_cfg3.json", "w"), indent=1)
assert res["slice28"]["rel_err"] < 1e-10 and res["slice31"]["rel_err"] < 1e-10
print("OK")
