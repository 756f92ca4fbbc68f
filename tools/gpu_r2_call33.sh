#!/bin/bash
# A/B after the cross-only rotations: CTAs per SM of the dataflow kernel, tail switch, on cfg 4 (chi = 512).
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
run() { env "$@" timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4', '$*', '->', round(d['value'],4), 'layers/s')"; }
run QTN_JACOBI_FLOW_CTAS=3
run QTN_JACOBI_FLOW_CTAS=4
run QTN_JACOBI_FLOW_CTAS=2
run QTN_JACOBI_FLOW_CTAS=3 QTN_JACOBI_TAIL=16
run QTN_JACOBI_FLOW_CTAS=4 QTN_JACOBI_TAIL=16
run QTN_JACOBI_FLOW_CTAS=3 QTN_JACOBI_INNER=2
