timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c 1-200
