mkdir -p gpurun_out
timeout 300 python tools/orth_time.py > gpurun_out/orth_time.log 2>&1
cat gpurun_out/orth_time.log
QTN_JACOBI_STATS=1 timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/bench_cfg5_cholqr.json 2> gpurun_out/bench_cfg5_cholqr.err
tail -1 gpurun_out/bench_cfg5_cholqr.json | cut -c1-300
grep -c . gpurun_out/bench_cfg5_cholqr.err; tail -5 gpurun_out/bench_cfg5_cholqr.err
