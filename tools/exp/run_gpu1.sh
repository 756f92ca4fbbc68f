set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_svd.py -x -q -m gpu -k "orth or mpo or rank_deficient or layer or mps" 2>&1 | tail -15 > gpurun_out/t_orth.log
cat gpurun_out/t_orth.log
timeout 300 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg5_cholqr.json 2> gpurun_out/bench_cfg5_cholqr.err
tail -1 gpurun_out/bench_cfg5_cholqr.json | cut -c1-400
tail -3 gpurun_out/bench_cfg5_cholqr.err
