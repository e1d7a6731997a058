timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for v in "" "DMG_NO_EARLY_KV=1"; do
env $v timeout 600 python bench.py --steps 1024 --warmup 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'])"
done
