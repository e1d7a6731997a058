# ncu evidence for the bench command (run only after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 12 --warmup 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 380 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel --launch-skip 40 --launch-count 2 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernelILi32 --launch-skip 20 --launch-count 4 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_attn.log gpurun_out/ncu_gemm.log
