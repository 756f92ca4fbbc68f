#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
QTN_JACOBI_TRACE=gpurun_out/trace_flow.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_flow.bin > gpurun_out/trace_cfg4_flow.txt; cat gpurun_out/trace_cfg4_flow.txt | head -60
rm -f gpurun_out/trace_flow.bin
