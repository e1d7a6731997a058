mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "bert or mask" 2>&1 | tail -40 > gpurun_out/c4p_pytest.log; tail -3 gpurun_out/c4p_pytest.log
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/c4p.json 2> gpurun_out/c4p.err; tail -1 gpurun_out/c4p.json | cut -c 1-200
timeout 300 python scripts/probe_bert_tc.py > gpurun_out/r2h_bert_tc_timeline.txt 2>&1; sed -n 14,28p gpurun_out/r2h_bert_tc_timeline.txt
