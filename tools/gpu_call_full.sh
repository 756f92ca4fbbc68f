set -x
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -25
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"zgemm_gather_kernel<\(int\)128, \(int\)16" -s 8 -c 10 -o gpurun_out/ncu_skinny_cfg3_search python bench.py --order search --max-log2 28 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
tail -3 gpurun_out/ncu_b2.log | cut -c1-300
ls -la gpurun_out | tail -4
