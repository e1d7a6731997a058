mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -rA --timeout 300 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/t_kernels.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -rA -s --timeout 600 > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/t_parity.log
timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
DMG_NO_SPLITK=1 timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline > gpurun_out/bench_nosplitk.log 2>&1
DMG_DECODE_V1=1 timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline > gpurun_out/bench_v1attn.log 2>&1
CMD="python bench.py --steps 12 --warmup 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_decode2 --launch-skip 40 --launch-count 2 -o gpurun_out/prof_attn2 -f $CMD > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_splitk --launch-skip 20 --launch-count 4 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
for f in t_kernels t_parity; do tail -n 4 gpurun_out/$f.log; done
tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-400
