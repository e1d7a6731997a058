mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2h_pytest.log; tail -2 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/c3u.json 2> gpurun_out/c3u.err; python -c "
import json; d=json.loads(open('gpurun_out/c3u.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['roofline']['frac'])"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/c4p.json 2> gpurun_out/c4p.err; python -c "
import json; d=json.loads(open('gpurun_out/c4p.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'])"
