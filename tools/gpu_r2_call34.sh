#!/bin/bash
# KP = 4 variant of the streaming kernel (K <= 4, N <= 8): parity (stream_check, contraction tests) and the notebook benchmark A/B.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 300 python tools/stream_check.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_contract.py tests/test_notebook_reference.py -m gpu -x -q 2>&1 | tail -3
for k in 0 1; do
QTN_STREAM_K4=$k timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); v=d['variants']; print('K4=$k', {a+'/'+b: round(v[a][b]['ms_device'],3) for a in v for b in ('default','optimized')}, {a+'/'+b: round(v[a][b]['ms_e2e_contract'],3) for a in v for b in ('default','optimized')})"
done
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-200 gpurun_out/bench_nbqft20_r02.json
timeout 900 python bench.py --order search --no-secondary --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('search', d['value'], d['roofline']['frac'])"
