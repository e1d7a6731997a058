mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err
tail -1 gpurun_out/bench_c2.log | cut -c 1-250; tail -3 gpurun_out/bench_c2.err
DMG_NO_GEMM_LN=1 timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline > gpurun_out/bench_c2_nofuse.log 2>&1
tail -1 gpurun_out/bench_c2_nofuse.log | cut -c 1-250
DMG_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/ncu_c3.log
