#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-600 gpurun_out/bench_nbqft20_r02.json; tail -3 gpurun_out/bench_nbqft20_r02.err
