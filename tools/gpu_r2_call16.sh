#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
echo "--- flow, 24 x 1024^2 (V accumulated)"
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | tail -5 | cut -c1-900
echo "--- old, 24 x 1024^2"
QTN_JACOBI_FLOW=0 QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | tail -5 | cut -c1-900
echo "--- flow, mixed sizes 12 x 600x520"
QTN_JACOBI_FLOW=2 timeout 300 python tools/svd_time.py 12 600 520 1 2>&1 | tail -2
QTN_JACOBI_FLOW=2 timeout 300 python tools/svd_time.py 3 333 700 1 2>&1 | tail -2
echo "--- cfg4 flow"
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 flow ->', round(d['value'],4), 'layers/s')"
QTN_JACOBI_FLOW=0 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 old ->', round(d['value'],4), 'layers/s')"
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q 2>&1 | tail -3
