#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (small problems): memcheck + racecheck of the dataflow Jacobi kernel
# (forced, mixed shapes) and of the streaming GEMM kernel.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
cat > /tmp/san_stream.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
q = g.load_package()
rng = np.random.default_rng(3)
for (M, N, K) in ((16384 + 5, 16, 16), (16384, 8, 8), (20000, 32, 20), (16400, 100, 32)):
    a = np.asfortranarray(rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K)))
    b = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
    got = q.ncon([a, b], [[-1, 1], [1, -2]])
    assert np.abs(got - a @ b).max() < 1e-11 * np.abs(a @ b).max()
    ak = np.asfortranarray(a.T)
    got = q.ncon([ak, b], [[1, -1], [1, -2]])
    assert np.abs(got - a @ b).max() < 1e-11 * np.abs(a @ b).max()
print("stream ok")
PY
for tool in memcheck racecheck; do
  echo "=== $tool: dataflow Jacobi (forced), 3 x 333x700 and 5 x 130x90"
  QTN_JACOBI_FLOW=2 timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python tools/svd_time.py 3 333 700 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max deviation|Error|hazard" | head -8
  QTN_JACOBI_FLOW=2 timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python tools/svd_time.py 5 130 90 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max deviation|Error|hazard" | head -8
  echo "=== $tool: streaming GEMM"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_stream.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|stream ok|Error|hazard" | head -8
done
