mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline"
for m in 0 1 2 4 8 16 32 64 72 52 127; do
  DMG_DEBUG_SKIP=$m $B > gpurun_out/skip_$m.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/skip_$m.log') if x.startswith('{')]
print('skip $m', 'ms/step %.4f' % json.loads(l[-1])['ms_per_step'] if l else 'FAILED')
PY
done
DMG_NO_PDL=1 DMG_DEBUG_SKIP=127 $B > gpurun_out/skip_127_nopdl.log 2>&1
DMG_NO_GRAPH=1 $B > gpurun_out/nograph.log 2>&1
python - <<PY
import json
for f in ['skip_127_nopdl','nograph']:
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    print(f, 'ms/step %.4f' % json.loads(l[-1])['ms_per_step'] if l else 'FAILED')
PY
