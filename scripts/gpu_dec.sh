mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decode or fused or generate or predict" 2>&1 | tail -4
timeout 600 python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/dec_u.json 2> gpurun_out/dec_u.err; python -c "
import json; d=json.loads(open('gpurun_out/dec_u.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['step_frac'])"
DMG_DECODE_TIMELINE=1 timeout 300 python scripts/probe_decode_layer.py > gpurun_out/dec_u_timeline.txt 2>&1; tail -32 gpurun_out/dec_u_timeline.txt
