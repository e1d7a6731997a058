timeout 300 python scripts/bench_attn_train.py
DMG_ATTN_FWD_MMA_SYNC=1 timeout 300 python scripts/bench_attn_train.py
