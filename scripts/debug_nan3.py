import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200.app_utils import baseline_config
from deepmusicgeneration_b200.model import get_language_model
from deepmusicgeneration_b200.training import TXLTrainer
from bench_train import lakh_shaped_tokens
B, T = 32, 512
L = int(os.environ.get('DBG_L', 16))
cfg = dict(baseline_config(), mask_steps=1, n_layers=L)
model = get_language_model(324, cfg, dtype='bf16', device=0, max_batch=1, max_seq=64, max_rows=64, keep_hidden=False, seed=0)
tr = TXLTrainer(model, B, T, cfg, drop_mult=float(os.environ.get('DBG_DROP', 1.0)), seed=7, distributed=False)
gen = torch.Generator().manual_seed(1234)
tok = lakh_shaped_tokens(B, 4 * T, gen)
xd = [tok[:, i * T:(i + 1) * T].contiguous().cuda() for i in range(4)]
yd = [tok[:, i * T + 1:(i + 1) * T + 1].contiguous().cuda() for i in range(4)]
tr.reset()
first = eval(os.environ.get('DBG_FIRST', '(1,0)'))
for s in range(3):
    mk = first if s == 0 else (1, 1)
    tr.forward(xd[s % 4], yd[s % 4], None, mask_size=mk); tr.backward()
    gr = tr.grads()
    bad = [k for k, v in gr.items() if not torch.isfinite(v).all()]
    big = sorted(((v.abs().max().item(), k) for k, v in gr.items()), reverse=True)[:3]
    tr.optimizer_step(float(os.environ.get('DBG_LR', 1e-4)))
    sd = model.state_dict()
    badw = [k for k, v in sd.items() if not torch.isfinite(v).all()]
    bigw = sorted(((v.abs().max().item(), k) for k, v in sd.items()), reverse=True)[:3]
    l = tr.losses()
    print(f'step {s} mask {mk} ce {l["ce"]:.4f} gnorm {l["grad_norm"]:.3f} bad grads {bad[:3]} max grads {big} bad weights {badw[:3]} max w {bigw}')
