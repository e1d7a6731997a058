mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py -x -q 2>&1 | tail -15 > gpurun_out/t_train.log
tail -6 gpurun_out/t_train.log
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
tail -2 gpurun_out/bench_c3.log | cut -c 1-200; tail -5 gpurun_out/bench_c3.err
DMG_ATTN_BWD_RECOMPUTE=1 timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_recompute.log 2>&1
tail -1 gpurun_out/bench_c3_recompute.log | cut -c 1-200
DMG_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/ncu_c3.log
