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


# ------------------------------------------------------------------------------------------------ data-parallel host logic (gloo, 2 ranks)
def test_bucket_plan_covers_every_layer_once_and_ends_small():
    from deepmusicgeneration_b200.training import TXLTrainer

    class Plan(TXLTrainer):
        def __init__(self, L, distributed, bucket_layers=4):
            self.n_layers, self.distributed, self.bucket_layers = L, distributed, bucket_layers

    for L in (1, 2, 3, 6, 8, 16, 24):
        assert Plan(L, False)._bucket_plan() == [(L, 0)]
        plan = Plan(L, True)._bucket_plan()
        assert plan[0][0] == L and plan[-1][1] == 0
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:])) and all(hi > lo for hi, lo in plan)
        assert plan[-1][0] - plan[-1][1] <= 2                       # the exposed bucket
        assert plan[0][0] - plan[0][1] <= 2                         # the first one leaves early


def _dp_worker(rank, world, port, q):
    import os
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    from deepmusicgeneration_b200 import sharding
    sharding.init_distributed(backend='gloo')
    torch.set_num_threads(2)
    cfg = dict(txl.default_config(), n_layers=2, d_model=32, n_heads=2, d_head=16, d_inner=64, mem_len=8, encode_position=False)
    # train mode draws rand_window_mask (a k = 0 window with p = 0.2 even at mask_steps = 1): pin every model of this test to (1, 1)
    txl.rand_window_mask = lambda x_len, m_len, device, **kw: txl.window_mask(x_len, device, m_len, size=(1, 1))
    torch.manual_seed(0)                                             # same weights on every rank
    m = txl.get_language_model(50, cfg, drop_mult=0.).train()
    m.reset()
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 50, (4, 8), generator=g); y = torch.randint(0, 50, (4, 8), generator=g)
    lo, hi = sharding.shard_range(4, rank, world)
    total, *_ = otrain.rnn_trainer_loss(m(x[lo:hi]), y[lo:hi])
    total.backward()
    params = otrain.unique_params(m)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    wire = flat.to(torch.bfloat16)                                   # the exchange's wire format (dmg_train_grad_pack)
    dist.all_reduce(wire)                                            # SUM, like TXLTrainer._exchange
    flat32 = flat.clone(); dist.all_reduce(flat32)
    q.put((rank, (wire.float() / world).tolist(), (flat32 / world).tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_average_equals_full_batch_gradient():
    """The data-parallel contract of TXLTrainer: SUM all-reduce of the per-rank gradients, times 1/world inside Adam, is the gradient
    of the reference step on the concatenated batch (equal shards: every loss term is a mean); the bf16 wire format keeps it within
    the bf16 tolerance."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    cfg = dict(txl.default_config(), n_layers=2, d_model=32, n_heads=2, d_head=16, d_inner=64, mem_len=8, encode_position=False)
    torch.manual_seed(0)
    m = txl.get_language_model(50, cfg, drop_mult=0.).train()
    m.reset()
    saved_mask = txl.rand_window_mask
    txl.rand_window_mask = lambda x_len, m_len, device, **kw: txl.window_mask(x_len, device, m_len, size=(1, 1))
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 50, (4, 8), generator=g); y = torch.randint(0, 50, (4, 8), generator=g)
    try:
        total, *_ = otrain.rnn_trainer_loss(m(x), y)
    finally:
        txl.rand_window_mask = saved_mask
    total.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in otrain.unique_params(m)])
    for rank, wire_avg, f32_avg in res:
        f32_avg, wire_avg = torch.tensor(f32_avg), torch.tensor(wire_avg)
        assert ((f32_avg - ref).norm() / ref.norm()).item() < 1e-5
        assert ((wire_avg - ref).norm() / ref.norm()).item() < 1e-2
