mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/orth_launches.csv python tools/orth_time.py 3072 1536 cholonly > gpurun_out/orth_ncu.log 2>&1
tail -3 gpurun_out/orth_ncu.log
wc -l gpurun_out/orth_launches.csv
