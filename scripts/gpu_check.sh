# one GPU: the whole -m gpu suite, the C3 / C4 legs alone, the decode timeline (used for every kernel change of round 2h)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --no-extra-legs --no-cpu-baseline | tail -1 | cut -c 1-160
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 | tail -1 | cut -c 1-160
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 | tail -1 | cut -c 1-160
DMG_DECODE_TIMELINE=1 timeout 300 python scripts/probe_decode_layer.py | tail -32
timeout 300 python scripts/probe_bert_tc.py | tail -20
