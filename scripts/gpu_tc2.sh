mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_tc.log 2> gpurun_out/bench_c3_tc.err
tail -1 gpurun_out/bench_c3_tc.log | cut -c 1-200; tail -3 gpurun_out/bench_c3_tc.err
DMG_ATTN_FWD_MMA_SYNC=1 timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c 1-200
DMG_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_tc.csv python bench.py --workload c3 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/ncu_c3.log
