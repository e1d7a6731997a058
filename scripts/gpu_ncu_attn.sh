mkdir -p gpurun_out
DMG_BENCH_PROFILE=1 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:attn_train_fwd_kernel --launch-skip 20 -c 1 -o gpurun_out/prof_attn_train_fwd -f python bench.py --workload c3 > gpurun_out/ncu_full1.log 2>&1
tail -2 gpurun_out/ncu_full1.log
DMG_BENCH_PROFILE=1 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_dq_kernel --launch-skip 20 -c 1 -o gpurun_out/prof_attn_bwd_dq -f python bench.py --workload c3 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
DMG_BENCH_PROFILE=1 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:gemm_train_kernel --launch-skip 300 -c 12 -o gpurun_out/prof_gemm_train -f python bench.py --workload c3 > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
