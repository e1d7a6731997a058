mkdir -p gpurun_out
N=${N:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 512 --warmup 8 > gpurun_out/bench_c2_g$N.log 2> gpurun_out/bench_c2_g$N.err
tail -1 gpurun_out/bench_c2_g$N.log | cut -c 1-260; tail -2 gpurun_out/bench_c2_g$N.err | cut -c 1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --workload c3 --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c3_g$N.log 2> gpurun_out/bench_c3_g$N.err
tail -1 gpurun_out/bench_c3_g$N.log | cut -c 1-260; tail -2 gpurun_out/bench_c3_g$N.err | cut -c 1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c 1-200
