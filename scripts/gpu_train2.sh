mkdir -p gpurun_out
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
tail -3 gpurun_out/bench_c3.log; tail -5 gpurun_out/bench_c3.err
DMG_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
wc -l gpurun_out/launches_c3.csv
