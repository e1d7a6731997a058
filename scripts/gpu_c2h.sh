mkdir -p gpurun_out
DMG_KV_L2_PREFETCH=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pf.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pf.log 2>&1
tail -1 gpurun_out/ncu_pf.log | cut -c 1-80
