mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_dq_kernel --launch-skip 16 -c 1 -o gpurun_out/prof_dq_ps -f python scripts/bench_attn_train.py > gpurun_out/ncu_dq_ps.log 2>&1
tail -2 gpurun_out/ncu_dq_ps.log
