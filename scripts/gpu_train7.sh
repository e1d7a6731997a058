mkdir -p gpurun_out
DMG_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
