#!/bin/bash
# ncu --set full: the dataflow Jacobi kernel with cross-only rotations (one cached-Gram sweep of cfg 4), the single-matrix
# eigensolve with cross-only rotations, and the streaming / skinny GEMM launches of the notebook benchmark (L2-resident).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_flow --launch-skip 7 -c 1 -o gpurun_out/ncu_jacobi_flow_xrot_r02 -f python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_flow.log 2>&1; echo "ncu flow rc=$?"
python tools/ncu_summary.py gpurun_out/ncu_jacobi_flow_xrot_r02.ncu-rep > gpurun_out/ncu_jacobi_flow_xrot_r02.txt; head -30 gpurun_out/ncu_jacobi_flow_xrot_r02.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_eig --launch-skip 400 -c 2 -o gpurun_out/ncu_jacobi_eig_xrot_r02 -f python tools/svd_time.py 1 1536 1024 1 > gpurun_out/ncu_eig.log 2>&1; echo "ncu eig rc=$?"
python tools/ncu_summary.py gpurun_out/ncu_jacobi_eig_xrot_r02.ncu-rep > gpurun_out/ncu_jacobi_eig_xrot_r02.txt; head -24 gpurun_out/ncu_jacobi_eig_xrot_r02.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm_stream --launch-skip 200 -c 4 -o gpurun_out/ncu_nbqft20_stream_r02 -f python tools/nb_once.py > gpurun_out/ncu_nb.log 2>&1; echo "ncu nb rc=$?"
python tools/ncu_summary.py gpurun_out/ncu_nbqft20_stream_r02.ncu-rep > gpurun_out/ncu_nbqft20_stream_r02.txt; head -50 gpurun_out/ncu_nbqft20_stream_r02.txt
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out/*.ncu-rep
