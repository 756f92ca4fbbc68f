#!/bin/bash
# After the one-shot host-path work: profile, notebook benchmark line, full GPU suite, smoke, default bench line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
python tools/oneshot_profile.py 2>&1 | tail -6 | tee gpurun_out/oneshot_profile_r02.txt
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-300 gpurun_out/bench_nbqft20_r02.json; tail -3 gpurun_out/bench_nbqft20_r02.err
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r02_late.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default_r02b.json 2> gpurun_out/bench_default_r02b.err; cut -c1-200 gpurun_out/bench_default_r02b.json; tail -2 gpurun_out/bench_default_r02b.err
