mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -1 gpurun_out/bench_cfg5.json | cut -c1-250
timeout 300 python bench.py --workload cfg2 --steps 200 --warmup 10 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -1 gpurun_out/bench_cfg2.json | cut -c1-250
timeout 300 python bench.py --workload cfg4 --chi 256 --steps 3 --warmup 3 > gpurun_out/bench_cfg4_chi256.json 2> gpurun_out/bench_cfg4_chi256.err; tail -1 gpurun_out/bench_cfg4_chi256.json | cut -c1-250
