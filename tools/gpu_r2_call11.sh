#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for cfg in "4 0" "4 0.3" "4 0.6" "4 1.0" "3 0.5" "3 1.0" "2 1.0"; do
  set -- $cfg
  export QTN_JACOBI_GROUPS=$1 QTN_JACOBI_SKEW=$2
  timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups $1 skew $2 cfg4 ->', round(d['value'],4), 'layers/s')"
done
export QTN_JACOBI_GROUPS=4 QTN_JACOBI_SKEW=0.6
QTN_JACOBI_TRACE=gpurun_out/trace_cfg4.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_cfg4.bin > gpurun_out/trace_cfg4_skew.txt; sed -n 1,8p gpurun_out/trace_cfg4_skew.txt; sed -n 30,42p gpurun_out/trace_cfg4_skew.txt; tail -3 gpurun_out/trace_cfg4_skew.txt
rm -f gpurun_out/trace_cfg4.bin
unset QTN_JACOBI_GROUPS QTN_JACOBI_SKEW
timeout 600 python -m pytest tests/test_gpu_svd.py tests/test_gpu_contract.py -x -q 2>&1 | tail -2
