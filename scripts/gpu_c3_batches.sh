for b in 16 64; do
timeout 900 python bench.py --workload c3 --train-batch $b --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c3_b$b.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_b$b.log').read())
print('batch $b', d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
done
