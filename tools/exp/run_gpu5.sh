mkdir -p gpurun_out
timeout 300 python tools/exp/sanity_orth.py > gpurun_out/sanity_orth.log 2>&1; cat gpurun_out/sanity_orth.log
timeout 900 python -m pytest tests/test_gpu_svd.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t_svd.log; cat gpurun_out/t_svd.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/exp/sanity_orth.py > gpurun_out/memcheck_orth.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/memcheck_orth.log
