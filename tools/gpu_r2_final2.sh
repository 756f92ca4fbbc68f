#!/bin/bash
# Final validation of round 2 (late): full GPU suite, smoke, notebook benchmark line, default line (cfg3 + secondary).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r02_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-200 gpurun_out/bench_nbqft20_r02.json; tail -2 gpurun_out/bench_nbqft20_r02.err
timeout 900 python bench.py > gpurun_out/bench_default_r02c.json 2> gpurun_out/bench_default_r02c.err; cut -c1-200 gpurun_out/bench_default_r02c.json; tail -2 gpurun_out/bench_default_r02c.err
