#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for c in 3 4 2; do
echo "--- flow ctas=$c, 24 x 1024^2 (V accumulated)"
QTN_JACOBI_FLOW_CTAS=$c QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | tail -3 | cut -c1-200
QTN_JACOBI_FLOW_CTAS=$c timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 flow ctas=$c ->', round(d['value'],4), 'layers/s')"
done
echo "--- flow, mixed sizes"
QTN_JACOBI_FLOW=2 timeout 300 python tools/svd_time.py 12 600 520 1 2>&1 | tail -2
QTN_JACOBI_FLOW=2 timeout 300 python tools/svd_time.py 3 333 700 1 2>&1 | tail -2
QTN_JACOBI_TRACE=gpurun_out/trace_flow.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_flow.bin > gpurun_out/trace_cfg4_flow3.txt; head -8 gpurun_out/trace_cfg4_flow3.txt; tail -3 gpurun_out/trace_cfg4_flow3.txt
rm -f gpurun_out/trace_flow.bin
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q 2>&1 | tail -3
