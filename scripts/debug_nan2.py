import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200.app_utils import baseline_config
from deepmusicgeneration_b200.model import get_language_model
from deepmusicgeneration_b200.training import TXLTrainer
from bench_train import lakh_shaped_tokens
B, T = 32, 512
cfg = dict(baseline_config(), mask_steps=1)
model = get_language_model(324, cfg, dtype='bf16', device=0, max_batch=1, max_seq=64, max_rows=64, keep_hidden=False, seed=0)
tr = TXLTrainer(model, B, T, cfg, drop_mult=1.0, seed=7, distributed=False)
np.random.seed(1234)
gen = torch.Generator().manual_seed(1234)
tok = lakh_shaped_tokens(B, 4 * T, gen)
xd = [tok[:, i * T:(i + 1) * T].contiguous().cuda() for i in range(4)]
yd = [tok[:, i * T + 1:(i + 1) * T + 1].contiguous().cuda() for i in range(4)]
sync = bool(os.environ.get('DBG_SYNC'))
fixed = os.environ.get('DBG_MASK')
tr.reset()
for s in range(int(os.environ.get('DBG_STEPS', 6))):
    mk = eval(fixed) if fixed else None
    if os.environ.get('DBG_FIRST10'): mk = (1, 0) if s == 0 else (1, 1)
    ms = tr.forward(xd[s % 4], yd[s % 4], None, mask_size=mk); tr.backward(); tr.optimizer_step(1e-4)
    if sync:
        l = tr.losses(); print('step', s, 'mask', ms, round(l['ce'], 5), round(l['grad_norm'], 4))
print('final', tr.losses())
