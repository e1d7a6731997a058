mkdir -p gpurun_out
DMG_BENCH_PROFILE=1 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:attn_train_fwd_tc --launch-skip 20 -c 1 -o gpurun_out/prof_attn_fwd_tc -f python bench.py --workload c3 > gpurun_out/ncu_full_tc.log 2>&1
tail -2 gpurun_out/ncu_full_tc.log
ls -la gpurun_out/*.ncu-rep
