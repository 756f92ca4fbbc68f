#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r2_c15.log
timeout 900 python bench.py --order search --steps 5 --warmup 2 --no-secondary --dump-steps gpurun_out/steps_cfg3_search_stream_r02.txt > gpurun_out/bench_cfg3_search_stream_r02.json 2> gpurun_out/bench_cfg3_search_stream_r02.err
cut -c1-300 gpurun_out/bench_cfg3_search_stream_r02.json
timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-secondary > gpurun_out/bench_cfg2_r02.json 2> gpurun_out/bench_cfg2_r02.err
cut -c1-300 gpurun_out/bench_cfg2_r02.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:zgemm_stream --launch-skip 16 -c 6 -o gpurun_out/ncu_stream_cfg3_search_r02 -f python bench.py --order search --steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-parity-check > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log | cut -c1-300
ls -la gpurun_out
