mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -6 > gpurun_out/t_par.log; tail -4 gpurun_out/t_par.log
for cfg in "1 4" "2 4" "2 2" "2 8" "4 4" "3 4"; do
  set -- $cfg
  DMG_DECODE_LANES=$1 DMG_LANE_STAGES=$2 timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline > gpurun_out/bench_lanes_$1_$2.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_lanes_$1_$2.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('lanes $1 stages $2: ms/step %.4f  tok/s %.0f  e2e %.0f  attn frac %.3f  step_frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac']))
else:
    print('lanes $1 stages $2 FAILED'); print(open('gpurun_out/bench_lanes_$1_$2.log').read()[-600:])
PY
done
