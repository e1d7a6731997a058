# gpurun --gpus N: N=${N:-4}; the default line under torchrun (C2 + the data-parallel `train` leg with its gradient exchange + `bert`)
N=${N:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r2h_bench_default_${N}gpu.json 2> gpurun_out/r2h_bench_default_${N}gpu.err
tail -1 gpurun_out/r2h_bench_default_${N}gpu.json | cut -c 1-200; tail -2 gpurun_out/r2h_bench_default_${N}gpu.err | cut -c 1-300
