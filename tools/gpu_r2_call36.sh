#!/bin/bash
# N = 4 predicate-free epilogue of the streaming kernel: parity and the notebook benchmark.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 300 python tools/stream_check.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_contract.py tests/test_notebook_reference.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; python -c "
import json; d=json.load(open('gpurun_out/bench_nbqft20_r02.json')); v=d['variants']; print({a+'/'+b: (round(v[a][b]['ms_device'],3), round(v[a][b]['ms_e2e_contract'],3)) for a in v for b in ('default','optimized')})"
