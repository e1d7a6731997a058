mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_step.py -x -q -k "c5 or fused or fit_one_cycle or spans" 2>&1 | tail -15 > gpurun_out/t_new.log
tail -6 gpurun_out/t_new.log
timeout 900 python bench.py --workload c5 --steps 256 --warmup 8 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err
tail -1 gpurun_out/bench_c5.log | cut -c 1-1800; tail -3 gpurun_out/bench_c5.err
