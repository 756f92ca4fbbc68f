#!/bin/bash
# A/B: cross-only rotations for small batches on the three-kernel path (QTN_JACOBI_XROT_ALL); then the refreshed default line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for x in 0 1; do
echo "--- XROT_ALL=$x"
for shape in "4 512 512" "8 700 520" "2 2048 1024" "12 256 256"; do
QTN_JACOBI_XROT_ALL=$x QTN_JACOBI_STATS=1 timeout 300 python tools/svd_time.py $shape 3 2>&1 | grep -E "jacobi|max dev" | tail -2 | cut -c1-130
done
done
QTN_JACOBI_XROT_ALL=1 timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_size_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default_r02b.json 2> gpurun_out/bench_default_r02b.err; cut -c1-200 gpurun_out/bench_default_r02b.json; tail -2 gpurun_out/bench_default_r02b.err
