#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for P in 0 1; do for ST in 0 -1; do
  export QTN_JACOBI_EIGPRIO=$P
  if [ $ST = -1 ]; then unset QTN_JACOBI_STAGGER_US; else export QTN_JACOBI_STAGGER_US=$ST; fi
  timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('eigprio $P stagger $ST cfg4 ->', round(d['value'],4), 'layers/s')"
  QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | grep "jacobi" | tail -1 | cut -c1-100
done; done
unset QTN_JACOBI_STAGGER_US; unset QTN_JACOBI_EIGPRIO
QTN_JACOBI_TRACE=gpurun_out/trace_cfg4.bin timeout 600 python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline 2>/dev/null | cut -c1-120
python tools/jacobi_trace.py gpurun_out/trace_cfg4.bin > gpurun_out/trace_cfg4_prio.txt; sed -n 1,8p gpurun_out/trace_cfg4_prio.txt; sed -n 30,50p gpurun_out/trace_cfg4_prio.txt; tail -3 gpurun_out/trace_cfg4_prio.txt
rm -f gpurun_out/trace_cfg4.bin
timeout 600 python -m pytest tests/test_gpu_svd.py -x -q 2>&1 | tail -2
