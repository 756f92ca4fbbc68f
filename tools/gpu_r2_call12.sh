#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_r2_c12.log
timeout 900 python bench.py > gpurun_out/bench_default_r2_c12.json 2> gpurun_out/bench_default_r2_c12.err
cut -c1-1500 gpurun_out/bench_default_r2_c12.json
