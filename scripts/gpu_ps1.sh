timeout 900 python -m pytest tests/test_gpu_train_kernels.py -x -q -k "attention_train" 2>&1 | tail -15
timeout 300 python scripts/bench_attn_train.py
