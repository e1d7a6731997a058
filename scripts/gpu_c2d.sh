for v in "" "DMG_SPLITK_8=1"; do
env $v timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['ms_per_step'], d['value'])"
done
