#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
QTN_JACOBI_TRACE=gpurun_out/trace_svd24.bin QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 1 2>&1 | grep "jacobi\|deviation" | cut -c1-200
python tools/jacobi_trace.py gpurun_out/trace_svd24.bin > gpurun_out/trace_svd24.txt; head -12 gpurun_out/trace_svd24.txt
QTN_JACOBI_TRACE=gpurun_out/trace_cfg4.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_cfg4.bin > gpurun_out/trace_cfg4.txt; cat gpurun_out/trace_cfg4.txt | head -60
rm -f gpurun_out/trace_svd24.bin
