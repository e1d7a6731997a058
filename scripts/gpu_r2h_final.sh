# final evidence of round 2h: -m gpu suite, smoke, default bench line, ncu --set full of decode_dual_kernel and of the BERT attention,
# launch lists (decode step without the graph: ncu does not list the kernels of a replayed graph), timelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2h_pytest_gpu_tail.log; tail -2 gpurun_out/r2h_pytest_gpu_tail.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; tail -1 gpurun_out/r2h_bench_default.json | cut -c 1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bert_tc_kernel -s 12 -c 1 -o gpurun_out/r2h_bert_tc -f python bench.py --workload c4 --bert-batch 32 --steps 1 --warmup 3 > gpurun_out/ncu_bert_r2h.log 2>&1; ls -la gpurun_out/r2h_bert_tc.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_c4_forward.csv python bench.py --workload c4 --bert-batch 32 --steps 1 --warmup 3 > gpurun_out/ncu_c4_list.log 2>&1; tail -1 gpurun_out/ncu_c4_list.log | cut -c 1-100
timeout 300 python scripts/probe_bert_tc.py > gpurun_out/r2h_bert_tc_timeline.txt 2>&1; grep -c tile gpurun_out/r2h_bert_tc_timeline.txt
if [ -n "$WITH_DECODE_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_dual -s 60 -c 1 -o gpurun_out/r2h_decode_dual -f python bench.py --steps 6 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/ncu_dec_full.log 2>&1; ls -la gpurun_out/r2h_decode_dual.ncu-rep
DMG_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:decode_dual|decode_layer|attn_decode3|sample_kernel|gemm_tc_splitk|state_advance|embed_kernel" -c 600 --csv --log-file gpurun_out/r2h_launches_decode_step.csv python bench.py --steps 6 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/ncu_dec_list.log 2>&1; tail -1 gpurun_out/ncu_dec_list.log | cut -c 1-120
DMG_DECODE_TIMELINE=1 timeout 300 python scripts/probe_decode_layer.py > gpurun_out/r2h_decode_timeline.txt 2>&1
fi
