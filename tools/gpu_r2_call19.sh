#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:jacobi_flow --launch-skip 7 -c 1 -o gpurun_out/ncu_jacobi_flow_r02 -f python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_flow.log 2>&1
tail -3 gpurun_out/ncu_flow.log | cut -c1-200
ls -la gpurun_out
