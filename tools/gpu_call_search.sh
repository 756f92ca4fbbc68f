set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_contract.py -m gpu -x -q -k "not full_size and not 3m" 2>&1 | tail -5
for s in 1 0; do QTN_SKINNY=$s timeout 120 python tools/skinny_check.py 2>&1 | tail -3; done
timeout 300 python bench.py --order search --steps 3 --warmup 3 --no-cpu-baseline --dump-steps gpurun_out/steps_cfg3_search_v2.txt > gpurun_out/bench_cfg3_search_v2.json 2> gpurun_out/bench_cfg3_search_v2.err
tail -c 700 gpurun_out/bench_cfg3_search_v2.json; tail -3 gpurun_out/bench_cfg3_search_v2.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg3_ref_v2.json 2> gpurun_out/bench_cfg3_ref_v2.err
tail -c 700 gpurun_out/bench_cfg3_ref_v2.json; tail -3 gpurun_out/bench_cfg3_ref_v2.err
timeout 200 python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2_ref_v2.json 2> gpurun_out/bench_cfg2_ref_v2.err; tail -c 300 gpurun_out/bench_cfg2_ref_v2.json
