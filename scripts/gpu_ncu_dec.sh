mkdir -p gpurun_out
DMG_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_splitk --launch-skip 2000 -c 4 -o gpurun_out/prof_dec_gemm -f python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_dec.log 2>&1
tail -2 gpurun_out/ncu_dec.log
