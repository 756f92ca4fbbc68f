#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for e in 128 256 512; do
echo "--- QTN_JACOBI_ET=$e single 1536x1024 / 4 x 512^2"
QTN_JACOBI_ET=$e QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 1 1536 1024 2 2>&1 | tail -3 | cut -c1-150
QTN_JACOBI_ET=$e QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 4 512 512 2 2>&1 | tail -3 | cut -c1-130
done
timeout 600 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 default ->', round(d['value'],4), d['unit'])"
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q 2>&1 | tail -3
