#!/bin/bash
# 2-GPU sanity of the default bench (as the driver launches it) after the plan-pool / plan-cache / Jacobi changes.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-secondary > gpurun_out/bench_default_n2_r02b.json 2> gpurun_out/bench_default_n2_r02b.err; cut -c1-300 gpurun_out/bench_default_n2_r02b.json; tail -3 gpurun_out/bench_default_n2_r02b.err
