#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for c in 3 2; do
QTN_JACOBI_FLOW_CTAS=$c timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 flow ctas=$c ->', round(d['value'],4), 'layers/s')"
done
timeout 600 python bench.py --workload cfg4 --chi 256 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 chi256 ->', round(d['value'],4), 'layers/s')"
QTN_JACOBI_FLOW=0 timeout 600 python bench.py --workload cfg4 --chi 256 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 chi256 old ->', round(d['value'],4), 'layers/s')"
QTN_JACOBI_TRACE=gpurun_out/trace_flow.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-100
python tools/jacobi_trace.py gpurun_out/trace_flow.bin > gpurun_out/trace_cfg4_flow4.txt; head -6 gpurun_out/trace_cfg4_flow4.txt; tail -2 gpurun_out/trace_cfg4_flow4.txt
rm -f gpurun_out/trace_flow.bin
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q 2>&1 | tail -3
