mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -x -q -k "attention_train" 2>&1 | tail -25 > gpurun_out/t_tc.log
tail -25 gpurun_out/t_tc.log
