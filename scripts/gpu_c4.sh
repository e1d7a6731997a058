mkdir -p gpurun_out
timeout 600 python scripts/bench_c4.py > gpurun_out/bench_c4.log 2>&1; tail -2 gpurun_out/bench_c4.log
C4_BATCH=32 C4_REPS=1 DMG_NO_FLASH=1 timeout 600 python scripts/bench_c4.py > gpurun_out/bench_c4_noflash.log 2>&1; tail -2 gpurun_out/bench_c4_noflash.log
