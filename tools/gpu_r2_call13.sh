#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 600 python tools/stream_check.py 2>&1 | tail -5
QTN_STREAM=0 timeout 600 python tools/stream_check.py 2>&1 | tail -2
timeout 900 python bench.py --order search --steps 5 --warmup 2 --no-cpu-baseline --no-secondary --dump-steps gpurun_out/steps_cfg3_search_stream.txt > gpurun_out/bench_cfg3_search_stream.json 2> gpurun_out/bench_cfg3_search_stream.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg3_search_stream.json').read().strip().splitlines()[-1])
print('stream: value', d['value'], 'ms', d['ms_per_step'], json.dumps(d['roofline'])[:600])
PY
awk 'NR>1 && $6>1.0' gpurun_out/steps_cfg3_search_stream.txt | sort -k6 -n -r | head -16
for s in 2 4; do
QTN_STREAM_STAGES=$s timeout 900 python bench.py --order search --steps 3 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stages $s value', d['value'], d['ms_per_step'])"
done
QTN_STREAM_WARPS=4 timeout 900 python bench.py --order search --steps 3 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps 4 value', d['value'], d['ms_per_step'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zgemm_stream --launch-skip 4 -c 8 -o gpurun_out/ncu_stream_cfg3_search_r02 -f python bench.py --order search --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
