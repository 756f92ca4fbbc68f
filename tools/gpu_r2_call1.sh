#!/bin/bash
# round-2 GPU call 1: smoke, Jacobi A/B (split vs fused), GPU tests, cfg4 bench A/B, ncu of the three round kernels
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 > gpurun_out/svd_split.log 2>&1; echo "svd split rc=$?"; tail -4 gpurun_out/svd_split.log | cut -c1-400
QTN_JACOBI_STATS=1 QTN_JACOBI=fused timeout 300 python tools/svd_time.py 24 1024 1024 2 > gpurun_out/svd_fused.log 2>&1; echo "svd fused rc=$?"; tail -4 gpurun_out/svd_fused.log | cut -c1-400
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 1 1536 1024 2 > gpurun_out/svd1_split.log 2>&1; echo "svd1 split rc=$?"; tail -3 gpurun_out/svd1_split.log | cut -c1-300
QTN_JACOBI_STATS=1 QTN_JACOBI=fused timeout 300 python tools/svd_time.py 1 1536 1024 2 > gpurun_out/svd1_fused.log 2>&1; echo "svd1 fused rc=$?"; tail -3 gpurun_out/svd1_fused.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
QTN_JACOBI_STATS=1 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_split.json 2> gpurun_out/cfg4_split.err; echo "cfg4 split rc=$?"; cut -c1-300 gpurun_out/cfg4_split.json; tail -2 gpurun_out/cfg4_split.err | cut -c1-600
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_fused.json 2> gpurun_out/cfg4_fused.err; echo "cfg4 fused(env unset => split again, control) rc=$?"
QTN_JACOBI=fused timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_fused.json 2> gpurun_out/cfg4_fused.err; echo "cfg4 fused rc=$?"; cut -c1-300 gpurun_out/cfg4_fused.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jacobi_ --launch-skip 600 -c 6 -o gpurun_out/ncu_jacobi_split_r02 -f python tools/svd_time.py 24 1024 1024 1 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
