mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
tail -2 gpurun_out/bench_c3.log | cut -c 1-300; tail -5 gpurun_out/bench_c3.err
DMG_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
