#!/bin/bash
# A/B: cross-only rotations in the cached-diagonal visits of the dataflow kernel (QTN_JACOBI_XROT_FLOW), cfg 4.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for x in 0 1 0 1; do
QTN_JACOBI_XROT_FLOW=$x QTN_JACOBI_STATS=1 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/st.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 xrot_flow=$x ->', round(d['value'],4), 'layers/s')"; tail -1 gpurun_out/st.txt | cut -c1-160
done
QTN_JACOBI_XROT_FLOW=1 timeout 600 python bench.py --workload cfg4 --chi 256 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4 chi256 xrot_flow=1 ->', round(d['value'],4), 'layers/s')"
QTN_JACOBI_XROT_FLOW=1 timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -m gpu -x -q 2>&1 | tail -3
