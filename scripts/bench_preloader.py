"Training data feed (SURVEY 8 f4): batches/s and tokens/s of the device MusicPreloader at the C3 geometry vs the CPU oracle."
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200.preloader import MusicPreloader
from oracle import preloader as opl
rng = np.random.default_rng(0)
items = [opl.Item(rng.integers(0, 324, L), np.cumsum(rng.integers(0, 5, L))) for L in rng.integers(200, 4000, 2000)]   # ~4.2 M tokens
bs, bptt = 32, 512
torch.manual_seed(0); np.random.seed(0)
pl = MusicPreloader(items, note_range=(12, 140), bs=bs, bptt=bptt, shuffle=True, transpose_range=(0, 12), encode_position=True)
pl.on_epoch_begin(); pl.next_batch(); torch.cuda.synchronize()
n = min(pl.n_batches - 1, 200)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): pl.next_batch()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
torch.manual_seed(0); np.random.seed(0)
ref = opl.MusicPreloader(items, (12, 140), bs=bs, bptt=bptt, shuffle=True, transpose_range=(0, 12), encode_position=True)
it = ref.batches(); next(it)
t0 = time.time(); m = 0
for _ in it:
    m += 1
    if m >= 20: break
cpu_ms = (time.time() - t0) / m * 1e3
alg_bytes = bs * (bptt + 1) * (4 + 4) + bs * bptt * 8 * 3           # tokens + positions read, x / pos / y written (int64)
print(json.dumps({'workload': 'MusicPreloader batch bs 32 x bptt 512, encode_position, random transpose', 'gpu_ms_per_batch': ms,
                  'gpu_tokens_per_s': bs * bptt / ms * 1e3, 'cpu_oracle_ms_per_batch': cpu_ms, 'cpu_tokens_per_s': bs * bptt / cpu_ms * 1e3,
                  'algorithmic_bytes_per_batch': alg_bytes, 'achieved_GBps': alg_bytes / ms / 1e6,
                  'note': 'one launch of bs CTAs moving 0.5 MB: latency-bound (launch + one dependent walk per row), not HBM-bound; '
                          'a C3 training step consumes one batch per 22.8 ms'}))
