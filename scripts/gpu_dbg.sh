python scripts/debug_nan3.py 2>&1 | tail -4
DBG_L=2 python scripts/debug_nan3.py 2>&1 | tail -4
DBG_DROP=0 python scripts/debug_nan3.py 2>&1 | tail -4
