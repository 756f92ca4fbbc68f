set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_native_network.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --order search --steps 5 --warmup 3 --dump-steps gpurun_out/steps_cfg3_search_v4.txt > gpurun_out/bench_cfg3_search_v4.json 2> gpurun_out/bench_cfg3_search_v4.err
tail -c 1000 gpurun_out/bench_cfg3_search_v4.json; tail -3 gpurun_out/bench_cfg3_search_v4.err
for k in 10 16; do
timeout 300 python bench.py --order search --open-wires $k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_search_open$k.json 2> gpurun_out/bench_cfg3_search_open$k.err
head -c 300 gpurun_out/bench_cfg3_search_open$k.json; tail -3 gpurun_out/bench_cfg3_search_open$k.err
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_cfg3_search.csv python bench.py --order search --max-log2 28 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
wc -l gpurun_out/launches_cfg3_search.csv
