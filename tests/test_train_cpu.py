"""CPU checks of the training host logic and of the oracle's restatement of the fastai step (no GPU needed)."""
import ctypes as C
import math

import numpy as np
import torch

from deepmusicgeneration_b200 import _lib
from deepmusicgeneration_b200.training import one_cycle_lr, rand_window_mask_size
from oracle import train as otrain
from oracle import txl


def test_train_config_struct_matches_header():
    assert C.sizeof(_lib.TrainConfig) == 64
    assert _lib.TrainConfig.seed.offset == 40 and _lib.TrainConfig.alpha.offset == 28


def test_rand_window_mask_size_follows_reference_draws():
    "deep_music_genre.py:1586-1590: rand() >= p -> (1,1) else (randint(0, max_size) + 1, 0), same RNG call order"
    rng, ref = np.random.RandomState(0), np.random.RandomState(0)
    for _ in range(300):
        got = rand_window_mask_size(4, p=0.2, is_eval=False, rng=rng)
        exp = (1, 1) if ref.rand() >= 0.2 else (ref.randint(0, 4) + 1, 0)
        assert got == exp
    assert rand_window_mask_size(4, is_eval=True) == (1, 1)
    assert rand_window_mask_size(None) == (1, 1)


def test_one_cycle_schedule_shape():
    lrs = [one_cycle_lr(i, 100, 1e-3)[0] for i in range(100)]
    moms = [one_cycle_lr(i, 100, 1e-3)[1] for i in range(100)]
    assert abs(lrs[0] - 1e-3 / 25) < 1e-9 and abs(max(lrs) - 1e-3) < 1e-6 and lrs.index(max(lrs)) == 30
    assert lrs[-1] < 1e-5 and all(a <= b + 1e-12 for a, b in zip(lrs[:30], lrs[1:31]))
    assert abs(moms[0] - 0.95) < 1e-9 and abs(min(moms) - 0.85) < 1e-3


def test_oracle_adam_true_wd_equals_adamw():
    "fastai true_wd (p *= 1 - lr*wd, then Adam) is torch.optim.AdamW's update"
    torch.manual_seed(0)
    w0 = torch.randn(7, 5)
    a, b = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w0.clone())
    opt_a = otrain.AdamTrueWD([a], betas=(0.9, 0.99), eps=1e-8)
    opt_b = torch.optim.AdamW([b], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)
    for _ in range(5):
        g = torch.randn(7, 5)
        a.grad, b.grad = g.clone(), g.clone()
        opt_a.step(1e-2, wd=0.01, clip=None)
        opt_b.step()
    assert torch.allclose(a, b, atol=1e-6)


def test_oracle_train_step_loss_terms_and_memory():
    "RNNTrainer terms: AR on the last layer output, TAR on the (updated, detached) last memory; mems advance by bptt"
    cfg = dict(txl.default_config(), n_layers=2, d_model=32, n_heads=2, d_head=16, d_inner=64, mem_len=12, encode_position=False)
    torch.manual_seed(0)
    m = txl.get_language_model(50, cfg, drop_mult=0.).train()
    m.reset()
    opt = otrain.AdamTrueWD(otrain.unique_params(m))
    x = torch.randint(0, 50, (3, 8)); y = torch.randint(0, 50, (3, 8))
    out = m(x)
    total, ce, ar, tar = otrain.rnn_trainer_loss(out, y, alpha=2., beta=1.)
    core = out[2][-1]
    assert torch.allclose(ar, 2. * core.pow(2).mean())
    h = out[1][-1]
    assert h.shape == (3, 8, 32) and not h.requires_grad
    assert torch.allclose(tar, (h[:, 1:] - h[:, :-1]).pow(2).mean())
    assert math.isclose(total.item(), (ce + ar + tar).item(), rel_tol=1e-6)
    m.reset()
    r1 = otrain.train_step(m, x, y, opt, 1e-3)
    r2 = otrain.train_step(m, x, y, opt, 1e-3)
    assert m[0].hidden[0].shape[1] == 12 and r1['grad_norm'] > 0 and r2['loss'] != r1['loss']


def test_oracle_fixed_masks_replace_every_dropout():
    cfg = dict(txl.default_config(), n_layers=2, d_model=32, n_heads=2, d_head=16, d_inner=64, mem_len=0, encode_position=False)
    m = txl.get_language_model(50, cfg, drop_mult=1.).train()
    sites = otrain.install_dropout_masks(m)
    assert [len(sites[k]) for k in ('embed', 'attn', 'res1', 'ff', 'res2', 'out')] == [1, 2, 2, 2, 2, 1]
    assert not any(isinstance(mod, torch.nn.Dropout) or type(mod).__name__ == 'RNNDropout' for mod in m.modules())
    m.reset()
    x = torch.randint(0, 50, (2, 6))
    np.random.seed(3)               # rand_window_mask draws from np.random in train mode (deep_music_genre.py:1587)
    a = m(x)[0]
    m.reset()
    np.random.seed(3)
    b = m(x)[0]
    assert torch.equal(a, b)        # no other randomness left
