mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_all.log; tail -3 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c 1-240
timeout 900 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.log | cut -c 1-240
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_tc.log 2> gpurun_out/bench_c4_tc.err; tail -1 gpurun_out/bench_c4_tc.log | cut -c 1-240
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4_tc.csv python bench.py --workload c4 --bert-batch 32 --steps 1 --warmup 3 > gpurun_out/ncu_c4_tc.log 2>&1; tail -1 gpurun_out/ncu_c4_tc.log | cut -c 1-100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bert_tc_kernel -s 5 -c 1 -o gpurun_out/prof_bert_tc_h16 -f python bench.py --workload c4 --bert-batch 32 --steps 1 --warmup 3 > gpurun_out/ncu_bert_tc_h16_full.log 2>&1; ls -la gpurun_out/prof_bert_tc_h16*
