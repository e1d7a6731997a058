mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q -s 2>&1 | tail -60 > gpurun_out/t_train_step.log
tail -40 gpurun_out/t_train_step.log
