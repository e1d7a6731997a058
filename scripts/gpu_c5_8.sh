mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --workload c5 --gpus 8 --steps 128 --warmup 4 --no-cpu-baseline > gpurun_out/bench_c5_g8.log 2> gpurun_out/bench_c5_g8.err
tail -1 gpurun_out/bench_c5_g8.log | cut -c 1-300; tail -2 gpurun_out/bench_c5_g8.err | cut -c 1-200
