timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lanes" 2>&1 | tail -3
for cfg in "1 4 0" "2 4 0" "2 4 1" "2 8 0" "2 3 0" "3 4 0"; do
  set -- $cfg
  if [ "$3" = "1" ]; then export DMG_LANE_NO_STAGGER=1; else unset DMG_LANE_NO_STAGGER; fi
  DMG_DECODE_LANES=$1 DMG_LANE_STAGES=$2 timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes $1 stages $2 nostagger $3:', d['ms_per_step'], d['value'])"
done
