mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bert_tc_kernel -s 12 -c 1 -o gpurun_out/r2h_bert_tc -f python bench.py --workload c4 --bert-batch 32 --steps 1 --warmup 3 > gpurun_out/ncu_bert_r2h.log 2>&1; ls -la gpurun_out/r2h_bert_tc*
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_c4_forward.csv python bench.py --workload c4 --bert-batch 32 --steps 1 --warmup 3 > gpurun_out/ncu_c4_list.log 2>&1; tail -2 gpurun_out/ncu_c4_list.log | cut -c 1-200
