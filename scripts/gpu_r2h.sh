mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2h_pytest.log; tail -2 gpurun_out/r2h_pytest.log
DMG_DECODE_TIMELINE=1 timeout 300 python scripts/probe_decode_layer.py > gpurun_out/r2h_timeline.txt 2>&1; tail -40 gpurun_out/r2h_timeline.txt
timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; tail -1 gpurun_out/r2h_bench_default.json | cut -c 1-300
