#!/bin/bash
# A/B: tail sweeps of a dataflow batch on the three-kernel path (QTN_JACOBI_TAIL = divisor of pairs_per_sweep), cfg 4.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for t in 0 16 6 3; do
QTN_JACOBI_XROT_FLOW=1 QTN_JACOBI_TAIL=$t QTN_JACOBI_STATS=1 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/st_tail$t.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 tail=$t ->', round(d['value'],4), 'layers/s', d['config']['max_bond_after'])"; tail -1 gpurun_out/st_tail$t.txt | cut -c1-200
done
QTN_JACOBI_XROT_FLOW=1 QTN_JACOBI_TAIL=6 timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -m gpu -x -q 2>&1 | tail -3
