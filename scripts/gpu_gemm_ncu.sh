mkdir -p gpurun_out
GEMM_ONLY=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_train_kernel --launch-skip 4 -c 1 -o gpurun_out/prof_gemm_qkv -f python scripts/bench_gemm.py > gpurun_out/ncu_g1.log 2>&1
tail -3 gpurun_out/ncu_g1.log
GEMM_ONLY=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_train_kernel --launch-skip 10 -c 1 -o gpurun_out/prof_gemm_ff1 -f python scripts/bench_gemm.py > gpurun_out/ncu_g2.log 2>&1
tail -3 gpurun_out/ncu_g2.log
