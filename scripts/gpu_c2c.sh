mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
for bn in 0 32 64 128; do
DMG_SPLITK_BN=$bn timeout 600 python bench.py --steps 512 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('BN $bn', d['ms_per_step'], d['value'])"
done
