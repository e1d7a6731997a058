mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2h_pytest.log; tail -3 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/c3u.json 2> gpurun_out/c3u.err; tail -1 gpurun_out/c3u.json | cut -c 1-260
