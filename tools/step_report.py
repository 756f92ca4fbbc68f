"""Per-step device timing of one slice of a BASELINE workload (diagnostics, GPU box)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

q = graft.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
max_log2 = int(sys.argv[2]) if len(sys.argv) > 2 else 28
precision = sys.argv[3] if len(sys.argv) > 3 else "c128"
net = (q.circuits.cfg3_network() if wl == "cfg3" else q.circuits.cfg2_network())[0]
q.optimize_contraction_order(net)
il = q.contract_rep(net)
arrays = [t.data for t in net.tensors]
S = q.choose_slices([a.shape for a in arrays], il, None, max_log2, 1) if wl == "cfg3" else []
plan = q.ContractionPlan([a.shape for a in arrays], il, None, S, precision=precision)
plan.upload(arrays)
plan.time_steps(0)
ms = plan.time_steps(1 if plan.nslices > 1 else 0)
rows = []
for (M, N, K, fl), t in zip(plan.steps(), ms):
    rows.append(dict(M=M, N=N, K=K, invariant=fl & 1, kind=(fl >> 1) & 7, variant=(fl >> 4) & 15, split_k=fl >> 8, ms=t,
                     tflops=8.0 * M * N * K / (t * 1e-3) / 1e12, gbs=16.0 * (M * K + K * N + M * N) / (t * 1e-3) / 1e9))
tot = sum(r["ms"] for r in rows if not r["invariant"])
print("workload", wl, "slices", plan.nslices, "steps", len(rows), "per-slice ms", tot, "flops/slice", plan.flops_per_slice,
      "TFLOP/s", plan.flops_per_slice / tot / 1e9)
for r in sorted(rows, key=lambda r: -r["ms"])[:25]:
    print("M=2^%.1f N=2^%.1f K=2^%.1f inv=%d var=%d split=%d  %.3f ms  %.2f TF/s  %.0f GB/s" % (
        __import__("math").log2(r["M"]), __import__("math").log2(r["N"]), __import__("math").log2(r["K"]), r["invariant"],
        r["variant"], r["split_k"], r["ms"], r["tflops"], r["gbs"]))
print("DOMINANT_STEP_INDEX", max(range(len(rows)), key=lambda i: rows[i]["ms"]))
json.dump(rows, open(os.path.join("gpurun_out", "steps_%s.json" % wl), "w"))
