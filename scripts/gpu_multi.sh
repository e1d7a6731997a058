mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --workload c3 --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c3_g2.log 2> gpurun_out/bench_c3_g2.err
tail -1 gpurun_out/bench_c3_g2.log | cut -c 1-700; tail -3 gpurun_out/bench_c3_g2.err
timeout 900 python bench.py --workload c3 --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_g1.log 2> gpurun_out/bench_c3_g1.err
tail -1 gpurun_out/bench_c3_g1.log | cut -c 1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 512 --warmup 8 > gpurun_out/bench_c2_g2.log 2> gpurun_out/bench_c2_g2.err
tail -1 gpurun_out/bench_c2_g2.log | cut -c 1-300; tail -3 gpurun_out/bench_c2_g2.err
