set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_contract.py tests/test_golden.py -m gpu -x -q -k "not full_size and not 3m" 2>&1 | tail -5
for pp in 0 1; do QTN_PREPERMUTE=$pp timeout 200 python tools/slice_value.py 31 99 2>&1 | tail -1; done
for pp in 0 1; do QTN_PREPERMUTE=$pp timeout 200 python tools/slice_value.py 31 5 --search 2>&1 | tail -1; done
timeout 300 python bench.py --order search --steps 5 --warmup 3 --dump-steps gpurun_out/steps_cfg3_search_v3.txt > gpurun_out/bench_cfg3_search_v3.json 2> gpurun_out/bench_cfg3_search_v3.err
tail -c 900 gpurun_out/bench_cfg3_search_v3.json; tail -3 gpurun_out/bench_cfg3_search_v3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg3_search.csv python bench.py --order search --max-log2 28 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"zgemm_gather_kernel<128, 16" -c 12 -o gpurun_out/ncu_skinny_cfg3_search python bench.py --order search --max-log2 28 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out | tail -5
