mkdir -p gpurun_out
C4_BATCH=64 C4_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python scripts/bench_c4.py > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log | cut -c 1-200
