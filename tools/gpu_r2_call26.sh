#!/bin/bash
# Plan cache of qtn_contract: parity test, one-shot profile, notebook benchmark line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_contract.py tests/test_notebook_reference.py tests/test_c_harness.py tests/test_native_network.py -m gpu -x -q 2>&1 | tail -4
python tools/oneshot_profile.py 2>&1 | tail -6 | cut -c1-250
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-200 gpurun_out/bench_nbqft20_r02.json; tail -3 gpurun_out/bench_nbqft20_r02.err
QTN_PLAN_CACHE=0 timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nocache', {k: (v['plain'][k]['ms_e2e_contract'] if k!='whole' else v['plain'][k]['ms_e2e']) for v in [d['variants']] for k in ('default','optimized','whole')})"
