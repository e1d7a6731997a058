timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_gemm.py 2>&1 | head -14
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c 1-200
