#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py tests/test_gpu_contract.py -x -q 2>&1 | tail -3
for G in 1 2 4; do
  export QTN_JACOBI_GROUPS=$G
  timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups $G cfg4 ->', round(d['value'],4), 'layers/s')"
  QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | grep "jacobi" | tail -1 | cut -c1-100
done
unset QTN_JACOBI_GROUPS
QTN_JACOBI_TRACE=gpurun_out/trace_cfg4.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_cfg4.bin > gpurun_out/trace_cfg4_e2.txt; sed -n 1,8p gpurun_out/trace_cfg4_e2.txt; sed -n 30,42p gpurun_out/trace_cfg4_e2.txt; tail -3 gpurun_out/trace_cfg4_e2.txt
rm -f gpurun_out/trace_cfg4.bin
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 1 1536 1024 2 2>&1 | grep "jacobi" | tail -1 | cut -c1-100
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jacobi_ --launch-skip 2400 -c 24 --csv --log-file gpurun_out/launches_jacobi2.csv python tools/svd_time.py 24 1024 1024 1 > /dev/null 2>&1; grep -c jacobi gpurun_out/launches_jacobi2.csv
