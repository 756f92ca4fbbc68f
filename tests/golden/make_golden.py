"""Regenerates tests/golden/golden_r01.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself (Julia) cannot run in the build
environment and stores no numeric vectors, so these fixtures pin the ORACLE (and through it the
CUDA path) against regressions; the literal known answers of the reference's own tests are
asserted separately in tests/test_oracle_kats.py."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import circuits as oc_, contract as oc, network2graph as o2g, plan as op, mps_sim as osim, svd as osvd  # noqa: E402


def digest(xs):
    return hashlib.sha256(",".join(str(int(x)) for x in xs).encode()).hexdigest()


out = {}
for name, net in (("cfg2", oc_.cfg2_network()[0]), ("cfg3", oc_.cfg3_network()[0])):
    perm = o2g.optimize_contraction_order(net)
    il = oc.contract_rep(net)
    nodes, steps = op.contraction_tree(il)
    dims = op.label_dims([t.data for t in net.tensors], il)
    f, b, mx, _ = op.tree_cost(nodes, steps, dims)
    entry = {"perm_sha256": digest(perm), "perm_head": perm[:16], "flops": f, "bytes": b, "max_elems": mx, "steps": len(steps)}
    if name == "cfg2":
        entry["amplitude"] = [complex(oc.contract(net)).real, complex(oc.contract(net)).imag]
        entry["slices_2^12_min64"] = op.choose_slice_labels(nodes, steps, dims, 12, 64)
    else:
        entry["slices_2^28"] = op.choose_slice_labels(nodes, steps, dims, 28, 1)
    out[name] = entry
small = {}
for seed in (1, 2, 3):
    net, _, bits = oc_.cfg2_network(12, 10, seed=seed)
    o2g.optimize_contraction_order(net)
    v = complex(oc.contract(net))
    small[str(seed)] = {"bits": [int(b) for b in bits], "amplitude": [v.real, v.imag]}
out["cfg2_12q_d10"] = small
net, _, _ = oc_.cfg3_network(4, 4, 8, seed=21)
o2g.optimize_contraction_order(net)
v = complex(oc.contract(net))
out["cfg3_4x4_c8_seed21"] = {"amplitude": [v.real, v.imag]}
# MPS: discarded weights and bonds of a truncated brickwork run
rng = np.random.default_rng(12)
sites = osim.product_state(14)
disc = []
for layer in range(10):
    left = list(range(1 + layer % 2, 14, 2))
    disc += osim.apply_layer(sites, left, [oc_.haar_unitary(4, rng) for _ in left], 1e-10, 16)
out["mps_14_chi16"] = {"bonds": [s.shape[2] for s in sites], "disc_sum": float(np.sum(disc)), "norm2": osim.overlap(sites, sites).real}
s = np.exp(-np.arange(100.0))
out["truncation_rank_exp_spectrum"] = {str(er): osvd.truncation_rank(s, er) for er in (0.0, 1e-10, 1e-6, 1e-3, 0.5, 3.0)}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_r01.json"), "w"), indent=1)
print("written", {k: list(v)[:3] for k, v in out.items()})
