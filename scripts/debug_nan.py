import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200.app_utils import baseline_config
from deepmusicgeneration_b200.model import get_language_model
from deepmusicgeneration_b200.training import TXLTrainer
B = int(os.environ.get('DBG_B', 32)); T = int(os.environ.get('DBG_T', 512)); L = int(os.environ.get('DBG_L', 16))
drop = float(os.environ.get('DBG_DROP', 1.0))
cfg = dict(baseline_config(), mask_steps=1, n_layers=L, mem_len=int(os.environ.get('DBG_M', 512)))
model = get_language_model(324, cfg, dtype='bf16', device=0, max_batch=1, max_seq=64, max_rows=64, keep_hidden=False, seed=0)
tr = TXLTrainer(model, B, T, cfg, drop_mult=drop, seed=7, distributed=False)
g = torch.Generator().manual_seed(1)
tr.reset()
for s in range(4):
    x = torch.randint(12, 301, (B, T), generator=g); y = torch.randint(12, 301, (B, T), generator=g)
    mk = eval(os.environ.get('DBG_MASK', '(1, 1)'))
    tr.forward(x, y, None, mask_size=mk); tr.backward()
    l = tr.losses()
    bad = [n for n, gr in tr.grads().items() if not torch.isfinite(gr).all()]
    print(f'B={B} T={T} L={L} drop={drop} step {s}: ce {l["ce"]:.4f} ar {l["ar"]:.4f} tar {l["tar"]:.4f}; non-finite grads: {len(bad)} {bad[:4]}')
    tr.optimizer_step(1e-4)
