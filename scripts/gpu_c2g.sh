for P in 0 2 3 4 6 9; do
DMG_KV_L2_PREFETCH=$P timeout 600 python bench.py --steps 1024 --warmup 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('prefetch items $P:', d['ms_per_step'], d['value'])"
done
