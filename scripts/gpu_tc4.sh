timeout 600 python -m pytest tests/test_gpu_train_kernels.py -x -q -k "attention_train" 2>&1 | tail -3
timeout 300 python scripts/bench_attn_train.py
DMG_ATTN_FWD_MMA_SYNC=1 timeout 300 python scripts/bench_attn_train.py
