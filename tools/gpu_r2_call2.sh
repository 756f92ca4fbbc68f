#!/bin/bash
# round-2 GPU call 2: sub-batch streams A/B (1, 2, 4 groups), SVD tests, cfg4
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 
for G in 1 2 4; do
  QTN_JACOBI_GROUPS=$G QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 > gpurun_out/svd_g$G.log 2>&1; echo "svd groups=$G rc=$?"; grep -h "jacobi\|deviation" gpurun_out/svd_g$G.log | tail -2 | cut -c1-330
done
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 1 1536 1024 2 > gpurun_out/svd1.log 2>&1; echo "svd1 rc=$?"; tail -3 gpurun_out/svd1.log | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
for G in 1 2 4; do
  QTN_JACOBI_GROUPS=$G QTN_JACOBI_STATS=1 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_g$G.json 2> gpurun_out/cfg4_g$G.err; echo "cfg4 groups=$G rc=$?"; cut -c1-130 gpurun_out/cfg4_g$G.json; tail -1 gpurun_out/cfg4_g$G.err | cut -c1-400
done
