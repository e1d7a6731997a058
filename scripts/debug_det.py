import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200.app_utils import baseline_config
from deepmusicgeneration_b200.model import get_language_model
from deepmusicgeneration_b200.training import TXLTrainer
B, T, L = int(os.environ.get('DBG_B', 32)), 512, int(os.environ.get('DBG_L', 2))
cfg = dict(baseline_config(), mask_steps=1, n_layers=L)
g = torch.Generator().manual_seed(1)
xs = [torch.randint(12, 301, (B, T), generator=g).cuda() for _ in range(3)]
ys = [torch.randint(12, 301, (B, T), generator=g).cuda() for _ in range(3)]
def run():
    model = get_language_model(324, cfg, dtype='bf16', device=0, max_batch=1, max_seq=64, max_rows=64, keep_hidden=False, seed=0)
    tr = TXLTrainer(model, B, T, cfg, drop_mult=float(os.environ.get('DBG_DROP', 1.0)), seed=7, distributed=False)
    tr.reset()
    out = []
    for s in range(2):
        tr.forward(xs[s], ys[s], None, mask_size=(1, 1)); tr.backward()
        out.append(({k: v.clone() for k, v in tr.grads().items()}, tr.losses()))
        tr.optimizer_step(0.0)      # lr 0: weights unchanged, so step 1 differs from step 0 only by the memory
    tr.close()
    return out
a, b = run(), run()
for s in range(2):
    print('step', s, 'ce', a[s][1]['ce'], b[s][1]['ce'])
    rows = []
    for k in a[s][0]:
        d = (a[s][0][k] - b[s][0][k]).norm().item(); n = a[s][0][k].norm().item()
        rows.append((d / max(n, 1e-20), k, n))
    for r in sorted(rows, reverse=True)[:8]:
        print('   rel diff %.3e  %-40s norm %.4e' % r)
