"""Regenerates tests/golden/notebook_qft.json (run from the repo root: `python tests/golden/make_notebook_golden.py`).

Two kinds of entries:
* ``published``: literal numbers that the REFERENCE ITSELF produced -- cell outputs stored in its example notebooks
  (`examples/network2graph_example.ipynb` cells 4 and 6: `network_graph` of a 4-ring; the medians / allocation of
  `examples/expectation_value_optimization_example.ipynb`).  They are copied by hand, not computed here.
* ``qft*``: the closed network <random bond-2 MPS| qft_circuit(N) |same MPS> of that notebook (seeded restatement,
  oracle/circuits.py:notebook_expectation_network): structure counts, order digests, plan costs and the contracted
  value, all from the CPU oracle.  `bench.py --workload nbqft20` and the GPU tests compare against them."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import circuits as oc_, contract as oc, network2graph as o2g  # noqa: E402


def digest(xs):
    return hashlib.sha256(",".join(str(int(x)) for x in xs).encode()).hexdigest()


out = {"published": {
    # examples/network2graph_example.ipynb cell 4: `G, edge_idx = network_graph(TN)` for four 2x2 tensors in a ring
    "ring4_network_graph": {"nv": 4, "ne": 4, "edge_idx": {"1,2": [1], "2,3": [2], "1,4": [4], "3,4": [3]}},
    # examples/expectation_value_optimization_example.ipynb (medians, seconds; allocation in GiB)
    "qft20_default_s": 3.052, "qft20_default_alloc_gib": 6.90, "qft20_optimized_s": 0.2804, "qft20_whole_s": 0.9179,
    "qft20_decomposed_default_s": 28.279, "qft20_decomposed_optimized_s": 0.1746, "qft20_decomposed_whole_s": 0.7956}}
for N, dec in ((6, False), (10, False), (10, True), (20, False), (20, True)):
    net = oc_.notebook_expectation_network(N, is_decompose=dec)
    st = []
    v = complex(np.asarray(oc.contract(net, stats=st)).reshape(-1)[0])
    n2 = net.copy()
    perm = o2g.optimize_contraction_order(n2)
    st2 = []
    v2 = complex(np.asarray(oc.contract(n2, stats=st2)).reshape(-1)[0])
    assert abs(v - v2) <= 1e-10 * abs(v)
    cost = lambda s: [float(sum(8.0 * m * n * k for m, n, k in s)), float(sum(16.0 * (m * k + k * n + m * n) for m, n, k in s)),  # noqa: E731
                      int(max(m * n for m, n, k in s))]
    out["qft%d_%s" % (N, "decomposed" if dec else "plain")] = {
        "tensors": len(net.tensors), "contractions": len(net.contractions), "value": [v.real, v.imag],
        "default_steps": len(st), "default_cost": cost(st), "optimized_steps": len(st2), "optimized_cost": cost(st2),
        "perm_sha256": digest(perm), "perm_head": [int(p) for p in perm[:16]]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "notebook_qft.json"), "w"), indent=1)
print({k: (v if k == "published" else (v["tensors"], v["contractions"], v["value"])) for k, v in out.items()})
