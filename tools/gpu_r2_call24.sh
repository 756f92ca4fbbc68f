#!/bin/bash
# One-shot contract(net) host path: division-free offset tables, pooled plan device state, cached pinned staging.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
python tools/oneshot_profile.py 2>&1 | tail -6 | tee gpurun_out/oneshot_profile_r02.txt
timeout 900 python -m pytest tests/test_gpu_contract.py tests/test_notebook_reference.py tests/test_c_harness.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-300 gpurun_out/bench_nbqft20_r02.json; tail -3 gpurun_out/bench_nbqft20_r02.err
timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2', d['value'], d['e2e'])"
