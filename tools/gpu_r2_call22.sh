#!/bin/bash
# Round-2 late pass: notebook benchmark of the reference (nbqft20) -- GPU parity tests, bench lines, launch list --
# then the full GPU suite, smoke and the default bench line (with secondary cfg4 / cfg5 / nb_qft20).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_notebook_reference.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_notebook_r02.log
timeout 600 python bench.py --workload nbqft20 --steps 10 --warmup 3 > gpurun_out/bench_nbqft20_r02.json 2> gpurun_out/bench_nbqft20_r02.err; cut -c1-400 gpurun_out/bench_nbqft20_r02.json; tail -3 gpurun_out/bench_nbqft20_r02.err
timeout 300 python bench.py --impl reference --workload nbqft20 --steps 10 --warmup 1 > gpurun_out/bench_ref_nbqft20_r02.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_nbqft20_r02.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_nbqft20_r02.csv python tools/nb_once.py 2>&1 | tail -3
QTN_JACOBI_STATS=1 timeout 300 python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/jacobi_stats_cfg4_r02.txt | cut -c1-120; tail -2 gpurun_out/jacobi_stats_cfg4_r02.txt | cut -c1-1500
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r02_late.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default_r02b.json 2> gpurun_out/bench_default_r02b.err; cut -c1-200 gpurun_out/bench_default_r02b.json; tail -2 gpurun_out/bench_default_r02b.err
