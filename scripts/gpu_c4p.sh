mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bert" 2>&1 | tail -40 > gpurun_out/c4p_pytest.log; tail -3 gpurun_out/c4p_pytest.log
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/c4p.json 2> gpurun_out/c4p.err; tail -1 gpurun_out/c4p.json | cut -c 1-200
DMG_BERT_TC_ONE_ITEM=1 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/c4p_one.json 2> gpurun_out/c4p_one.err; tail -1 gpurun_out/c4p_one.json | cut -c 1-200
