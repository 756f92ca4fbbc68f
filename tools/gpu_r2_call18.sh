#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
echo "--- flow, 24 x 1024^2 (V accumulated)"
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | tail -3 | cut -c1-400
echo "--- flow, mixed sizes"
QTN_JACOBI_FLOW=2 timeout 300 python tools/svd_time.py 12 600 520 1 2>&1 | tail -2
QTN_JACOBI_FLOW=2 timeout 300 python tools/svd_time.py 3 333 700 1 2>&1 | tail -2
echo "--- cfg4 flow"
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 flow ->', round(d['value'],4), 'layers/s')"
QTN_JACOBI_TRACE=gpurun_out/trace_flow.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_flow.bin > gpurun_out/trace_cfg4_flow2.txt; head -8 gpurun_out/trace_cfg4_flow2.txt; tail -3 gpurun_out/trace_cfg4_flow2.txt
rm -f gpurun_out/trace_flow.bin
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q 2>&1 | tail -3
