#!/bin/bash
# round-2 GPU call 3: 3M update + cached-Gram mode A/B, SVD tests, cfg4, launch list
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for C in 0 1; do
  QTN_JACOBI_CROSS=$C QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 > gpurun_out/svd_c$C.log 2>&1; echo "svd cross=$C rc=$?"; grep -h "jacobi\|deviation" gpurun_out/svd_c$C.log | tail -2 | cut -c1-700
done
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 1 1536 1024 2 > gpurun_out/svd1.log 2>&1; echo "svd1 rc=$?"; tail -3 gpurun_out/svd1.log | cut -c1-400
QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 3 3072 1536 1 > gpurun_out/svd3.log 2>&1; echo "svd3 rc=$?"; tail -3 gpurun_out/svd3.log | cut -c1-400
timeout 1200 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
for C in 0 1; do
  QTN_JACOBI_CROSS=$C QTN_JACOBI_STATS=1 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/cfg4_c$C.json 2> gpurun_out/cfg4_c$C.err; echo "cfg4 cross=$C rc=$?"; cut -c1-130 gpurun_out/cfg4_c$C.json; tail -1 gpurun_out/cfg4_c$C.err | cut -c1-700
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jacobi_ --launch-skip 2400 -c 48 --csv --log-file gpurun_out/launches_jacobi.csv python tools/svd_time.py 24 1024 1024 1 > gpurun_out/ncu_l.log 2>&1; echo "ncu rc=$?"
