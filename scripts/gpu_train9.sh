mkdir -p gpurun_out
python scripts/debug_nan3.py 2>&1 | tail -3 | cut -c 1-400
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log
timeout 900 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
tail -1 gpurun_out/bench_c3.log | cut -c 1-2200; tail -5 gpurun_out/bench_c3.err
