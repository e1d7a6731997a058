mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -rA --timeout 300 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/t_kernels.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -rA -s --timeout 400 > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/t_parity.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 256 --warmup 8 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/t_kernels.log gpurun_out/t_parity.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench.err
