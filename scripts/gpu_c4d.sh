timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_c4.py 2>&1 | tail -1
