mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -x -q -k "persistent" 2>&1 | tail -12
for v in "" "DMG_GEMM_NO_AUX_TMA=1" "DMG_GEMM_NO_TMA_STORE=1" "DMG_GEMM_1CTA=1"; do
  env $v timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['config']['loss_before'], d['config']['loss_after'], d['ms_per_step'])"
done
