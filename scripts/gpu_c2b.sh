mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
echo "default bench took $SECONDS s"
tail -1 gpurun_out/bench_default.log | cut -c 1-2500; tail -3 gpurun_out/bench_default.err
