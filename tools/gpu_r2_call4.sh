#!/bin/bash
# round-2 GPU call 4: row-split / group tuning on cfg4, then the default bench (cfg3 + secondary cfg4/cfg5) and the reference arm
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for cfg in "4 2 2" "4 3 3" "4 4 4" "4 6 6" "4 4 2" "4 2 4" "3 4 4" "2 4 4" "4 8 8"; do
  set -- $cfg
  QTN_JACOBI_GROUPS=$1 QTN_JACOBI_S=$2 QTN_JACOBI_SU=$3 timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups S SU = $cfg ->', round(d['value'],4), 'layers/s')"
done
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
nproc; free -g | head -2
