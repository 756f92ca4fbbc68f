#!/bin/bash
# Round-2 validation pass: GPU test suite, smoke, default bench (cfg3 + secondary cfg4 / cfg5), reference arm,
# searched-order bench, cfg2 bench.  Outputs under gpurun_out/ (copied to profiles/ by hand).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r02_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default_r02.json 2> gpurun_out/bench_default_r02.err; cut -c1-200 gpurun_out/bench_default_r02.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; cut -c1-300 gpurun_out/bench_ref_r02.json
timeout 900 python bench.py --order search --no-secondary > gpurun_out/bench_cfg3_search_r02.json 2> gpurun_out/bench_cfg3_search_r02.err; cut -c1-200 gpurun_out/bench_cfg3_search_r02.json
timeout 600 python bench.py --workload cfg4 --chi 256 --steps 3 --warmup 2 > gpurun_out/bench_cfg4_chi256_r02.json 2>/dev/null; cut -c1-160 gpurun_out/bench_cfg4_chi256_r02.json
