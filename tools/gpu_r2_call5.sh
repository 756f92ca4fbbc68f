#!/bin/bash
# round-2 GPU call 5: update-kernel CTA size A/B (128 vs 256 threads) on cfg4 and the batched SVD
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for T in 128 256; do
  touch qaintensor.jl_b200/csrc/svd.cu
  QTN_NVCC_EXTRA="-DQTN_UPDATE_THREADS=$T" python -c "import __graft_entry__ as g; g.build()"
  for SU in 0 4 8; do
    if [ $SU = 0 ]; then unset QTN_JACOBI_SU; else export QTN_JACOBI_SU=$SU; fi
    timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('update threads $T SU=$SU cfg4 ->', round(d['value'],4), 'layers/s')"
  done
  unset QTN_JACOBI_SU
  QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py 24 1024 1024 2 2>&1 | grep "jacobi\|deviation" | tail -2 | cut -c1-120
done
touch qaintensor.jl_b200/csrc/svd.cu; python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests/test_gpu_svd.py -x -q 2>&1 | tail -2
