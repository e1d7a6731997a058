# gpurun --gpus 2: the default line under torchrun (C2 + the data-parallel `train` leg with its gradient exchange + `bert`) and the 2-rank tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2h_bench_default_2gpu.json 2> gpurun_out/r2h_bench_default_2gpu.err
tail -1 gpurun_out/r2h_bench_default_2gpu.json | cut -c 1-200; tail -2 gpurun_out/r2h_bench_default_2gpu.err | cut -c 1-300
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -3
