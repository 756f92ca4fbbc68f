#!/bin/bash
# A/B: cross-only rotations in the single-matrix eigensolve (QTN_JACOBI_XROT), cfg 5 and a single 1536 x 1024 SVD.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for x in 0 1; do
echo "--- XROT=$x"
QTN_JACOBI_XROT=$x QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 1 1536 1024 3 2>&1 | tail -3 | cut -c1-400
QTN_JACOBI_XROT=$x timeout 600 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 xrot=$x ->', round(d['value'],4), 'applies/s', d['config']['energy_after_last_apply'], d['config']['max_discarded_weight'])"
done
QTN_JACOBI_XROT=1 timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -m gpu -x -q 2>&1 | tail -3
