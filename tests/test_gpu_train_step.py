"""GPU parity of the whole training step (dmg_train_*) against the oracle's restatement of the fastai step
(oracle/train.py on oracle/txl.py): loss parts, every parameter gradient, the Adam update and the memory carried into the
next step, with and without dropout (the oracle is fed the very masks the kernels draw)."""
import numpy as np
import pytest
import torch

from deepmusicgeneration_b200.model import get_language_model
from deepmusicgeneration_b200.training import TXLTrainer, one_cycle_lr, rand_window_mask_size
from oracle import train as otrain
from oracle import txl

pytestmark = pytest.mark.gpu

V = 324


def small_config(**kw):
    c = dict(txl.default_config(), n_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256, mem_len=64, ctx_len=64,
             encode_position=False, mask_steps=1)
    c.update(kw)
    return c


def build_pair(cfg, bs, bptt, drop_mult, seed=0, alpha=2., beta=1.):
    torch.manual_seed(seed)
    om = txl.get_language_model(V, cfg, drop_mult=drop_mult)
    pm = get_language_model(V, cfg, dtype='bf16', device=0, max_batch=bs, max_seq=max(bptt, 64), keep_hidden=False, init=False)
    pm.load_state_dict(om.state_dict())
    tr = TXLTrainer(pm, bs, bptt, cfg, drop_mult=drop_mult, alpha=alpha, beta=beta, seed=1234, distributed=False)
    return om, pm, tr


def set_oracle_mask(monkeypatch, size):
    def fixed(x_len, m_len, device, max_size=None, p=0.2, is_eval=False, rng=None):
        return txl.window_mask(x_len, device, m_len, size=size)
    monkeypatch.setattr(txl, 'rand_window_mask', fixed)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def compare_grads(tr, om, tol, skip=()):
    got = tr.grads()
    ref = {n: p.grad for n, p in om.state_dict(keep_vars=True).items() if getattr(p, 'grad', None) is not None}
    checked = 0
    worst = (0., None)
    for name, g in got.items():
        if name in skip:
            continue
        r = ref.get(name)
        if r is None and name == '0.encoder.weight':
            r = ref.get('1.decoder.weight')
        assert r is not None, f'oracle has no gradient for {name}'
        e = rel(g, r.reshape(g.shape))
        if e > worst[0]:
            worst = (e, name)
        checked += 1
    assert worst[0] < tol, f'worst gradient mismatch {worst}'
    print('worst gradient rel err', worst)
    return checked


def install_masks(tr, sites, cfg, bs, bptt, m_len, step):
    "give the oracle the masks the kernels will draw for `step`"
    d, H, di, M = cfg['d_model'], cfg['n_heads'], cfg['d_inner'], cfg['mem_len']
    S = M + bptt
    sites['embed'][0].mask = tr.dropout_mask(0, 0, (bs, bptt, d), step).cpu()
    sites['out'][0].mask = tr.dropout_mask(5, 0, (bs, 1, d), step).cpu()
    for l in range(cfg['n_layers']):
        full = tr.dropout_mask(1, l, (bs, H, bptt, S), step)
        sites['attn'][l].mask = full[..., M - m_len:].cpu()
        sites['res1'][l].mask = tr.dropout_mask(2, l, (bs, bptt, d), step).cpu()
        sites['ff'][l].mask = tr.dropout_mask(3, l, (bs, bptt, di), step).cpu()
        sites['res2'][l].mask = tr.dropout_mask(4, l, (bs, bptt, d), step).cpu()


@pytest.mark.parametrize('bptt', [64, 128])      # 128 (with mem_len 128): the tcgen05 forward / saved-probability dQ / tcgen05 dK-dV path
@pytest.mark.parametrize('drop_mult,encode_position,mask_size', [(0., False, (1, 1)), (0., True, (1, 0)), (1., False, (1, 1)),
                                                                 (1., True, (2, 0))])
def test_training_step_matches_oracle(monkeypatch, drop_mult, encode_position, mask_size, bptt):
    cfg = small_config(encode_position=encode_position, mask_steps=2, mem_len=bptt, ctx_len=bptt)
    bs, steps = 3, 3
    om, pm, tr = build_pair(cfg, bs, bptt, drop_mult)
    om.train(); om.reset()
    sites = otrain.install_dropout_masks(om)
    set_oracle_mask(monkeypatch, mask_size)
    opt = otrain.AdamTrueWD(otrain.unique_params(om), eps=1e-3)
    sd0 = {k: v.clone() for k, v in om.state_dict().items()}
    tr.reset()
    g = torch.Generator().manual_seed(5)
    lr = 1e-3
    for s in range(steps):
        x = torch.randint(0, V, (bs, bptt), generator=g)
        y = torch.randint(0, V, (bs, bptt), generator=g)
        pos = torch.cumsum(torch.randint(0, 9, (bs, bptt), generator=g), 1) + 40 * s if encode_position else None
        m_len = min(cfg['mem_len'], s * bptt)
        if drop_mult > 0:
            install_masks(tr, sites, cfg, bs, bptt, m_len, s)
        ref = otrain.train_step(om, {'x': x, 'pos': pos} if encode_position else x, y, opt, lr, wd=0.01, clip=0.5)
        tr.forward(x, y, pos, mask_size=mask_size)
        tr.backward()
        got = tr.losses()
        assert abs(got['ce'] - ref['ce']) < 2e-2 * max(1., abs(ref['ce'])), (s, got, ref)
        assert abs(got['ar'] - ref['ar']) < 2e-2 * max(1e-3, abs(ref['ar'])), (s, got, ref)
        assert abs(got['tar'] - ref['tar']) < 3e-2 * max(1e-3, abs(ref['tar'])), (s, got, ref)
        n = compare_grads(tr, om, tol=2e-2)
        assert n >= 4 + 11 * cfg['n_layers']
        tr.optimizer_step(lr, betas=(0.9, 0.99), eps=1e-3, wd=0.01, clip=0.5)
        got = tr.losses()
        assert abs(got['grad_norm'] - ref['grad_norm']) < 3e-2 * ref['grad_norm'], (s, got, ref)
    # parameter movement after three Adam steps (eps is large in this test so that the update is a smooth function of the
    # gradient: with the default 1e-8 an update is lr*sign(g) and bf16 noise on near-zero gradients flips signs)
    sd_ref, sd_got = om.state_dict(), pm.state_dict()
    for name, w in sd_got.items():
        d_got, d_ref = w - sd0[name].reshape(w.shape), (sd_ref[name] - sd0[name]).reshape(w.shape)
        assert rel(d_got, d_ref) < 0.1, (name, rel(d_got, d_ref))
    tr.close()


def test_training_reduces_loss_and_inference_follows():
    "a few hundred steps on a repeating pattern: the loss must fall; afterwards the inference path sees the trained weights"
    cfg = small_config()
    bs, bptt = 4, 64
    om, pm, tr = build_pair(cfg, bs, bptt, drop_mult=1.0, alpha=2., beta=1.)
    base = torch.arange(bs * bptt * 40) % 37 + 12
    data = base.view(bs, -1)
    tr.reset()
    first = last = None
    n = data.shape[1] // bptt - 1
    for ep in range(3):
        tr.reset()
        for i in range(n):
            x = data[:, i * bptt:(i + 1) * bptt]
            y = data[:, i * bptt + 1:(i + 1) * bptt + 1]
            lr, mom = one_cycle_lr(ep * n + i, 3 * n, 3e-3)
            tr.step(x, y, lr=lr, betas=(mom, 0.99))
            if first is None:
                first = tr.losses()['ce']
    last = tr.losses()['ce']
    assert first > 4.0 and last < 0.5 * first, (first, last)
    tr.sync_for_inference()
    pm.reset()
    logits = pm(data[:, :bptt].cuda())[0]
    pred = logits.argmax(-1).cpu()
    acc = (pred[:, 8:] == data[:, 9:bptt + 1]).float().mean().item()
    assert acc > 0.9, acc
    tr.close()




def test_grad_spans_tile_the_flat_buffer():
    "backward slices (as the trainer issues them for the bucketed all-reduce) finalise disjoint spans that cover every gradient"
    import ctypes as C
    cfg = small_config(n_layers=5)
    om, pm, tr = build_pair(cfg, 2, 64, 0.)
    lib, h = tr.lib, tr.e.h
    total = lib.dmg_train_param_count(h)
    for bucket in (1, 2, 3, 5, 8):
        spans, hi = [], 5
        while True:
            lo = max(0, hi - bucket)
            off, cnt = C.c_int64(), C.c_int64()
            assert lib.dmg_train_grad_span(h, hi, lo, C.byref(off), C.byref(cnt)) == 0
            spans.append((off.value, cnt.value))
            hi = lo
            if hi == 0:
                break
        pos = 0
        for off, cnt in spans:
            assert off == pos and cnt > 0
            pos += cnt
        assert pos == total
    tr.close()


def test_learner_fit_one_cycle_surface():
    "music_model_learner(...).fit_one_cycle(epochs, lr, batches): the notebook's training call on the CUDA engine"
    from deepmusicgeneration_b200.codec import MusicDataBunch
    from deepmusicgeneration_b200.learner import music_model_learner
    cfg = small_config()
    data = MusicDataBunch.empty('')
    learn = music_model_learner(data, config=cfg, dtype='bf16', max_batch=4, max_seq=64, keep_hidden=False, seed=0)
    base = (torch.arange(4 * 64 * 12) % 29 + 12).view(4, -1)
    batches = [(base[:, i * 64:(i + 1) * 64], base[:, i * 64 + 1:(i + 1) * 64 + 1]) for i in range(11)]
    seen = []
    out = learn.fit_one_cycle(4, 3e-3, batches, callback=lambda i, tr: seen.append(i))
    assert len(seen) == 44 and out['ce'] < 2.5, out
    learn.model.reset()
    logits = learn.model(base[:, :64].cuda())[0]
    acc = (logits.argmax(-1).cpu()[:, 8:] == base[:, 9:65]).float().mean().item()
    assert acc > 0.9, acc


def test_training_memory_longer_than_segment(monkeypatch):
    "mem_len 128 > bptt 64: the memory grows 0 -> 64 -> 128 and then slides (copy path, not the buffer swap of bptt == mem_len)"
    cfg = small_config(mem_len=128)
    bs, bptt = 2, 64
    om, pm, tr = build_pair(cfg, bs, bptt, 0.)
    om.train(); om.reset()
    otrain.install_dropout_masks(om)
    set_oracle_mask(monkeypatch, (1, 1))
    opt = otrain.AdamTrueWD(otrain.unique_params(om), eps=1e-3)
    tr.reset()
    g = torch.Generator().manual_seed(9)
    for s in range(4):
        x = torch.randint(0, V, (bs, bptt), generator=g)
        y = torch.randint(0, V, (bs, bptt), generator=g)
        ref = otrain.train_step(om, x, y, opt, 1e-3, wd=0.01, clip=0.5)
        tr.forward(x, y, None, mask_size=(1, 1))
        tr.backward()
        got = tr.losses()
        assert abs(got['ce'] - ref['ce']) < 2e-2 * max(1., abs(ref['ce'])), (s, got, ref)
        assert abs(got['tar'] - ref['tar']) < 3e-2 * max(1e-3, abs(ref['tar'])), (s, got, ref)
        compare_grads(tr, om, tol=2e-2)
        tr.optimizer_step(1e-3, betas=(0.9, 0.99), eps=1e-3, wd=0.01, clip=0.5)
    tr.close()


def test_loss_curve_50_steps_c3_geometry(monkeypatch):
    """N-step loss-curve parity at the C3 width (d_model 512, 8 heads x 64, d_inner 2048, bptt = mem_len = 512: the tcgen05 attention
    forward / saved-probability backward / persistent GEMMs of the benchmark), 4 layers, 2 sequences, 50 steps of the notebook's recipe
    (Adam(0.9, 0.99), true_wd 0.01, clip 0.5, default eps) on four repeating LakhMIDI-shaped batches, dropout off so that both sides
    are deterministic: every loss part of every step follows the oracle's fastai step, and the loss falls."""
    import bench_train
    cfg = dict(txl.baseline_config(), n_layers=4, mask_steps=1)
    bs, bptt, steps = 2, 512, 50
    om, pm, tr = build_pair(cfg, bs, bptt, 0.)
    om.train(); om.reset()
    set_oracle_mask(monkeypatch, (1, 1))
    opt = otrain.AdamTrueWD(otrain.unique_params(om))
    toks = bench_train.lakh_shaped_tokens(bs, 4 * bptt, torch.Generator().manual_seed(3))
    tr.reset()
    lr = 5e-4
    curve = []
    for s in range(steps):
        i = s % 4
        x, y = toks[:, i * bptt:(i + 1) * bptt], toks[:, i * bptt + 1:(i + 1) * bptt + 1]
        ref = otrain.train_step(om, x, y, opt, lr, wd=0.01, clip=0.5)
        tr.step(x, y, lr=lr, mask_size=(1, 1))
        got = tr.losses()
        curve.append((ref['ce'], got['ce'], ref['ar'], got['ar'], ref['tar'], got['tar']))
    for s, (rce, gce, rar, gar, rtar, gtar) in enumerate(curve):
        assert abs(gce - rce) < 3e-2 * rce + 2e-2, (s, curve[s])
        assert abs(gar - rar) < 3e-2 * rar + 1e-3, (s, curve[s])
        assert abs(gtar - rtar) < 5e-2 * rtar + 1e-3, (s, curve[s])
    print('loss curve (oracle ce, cuda ce) every 10 steps:', [(round(c[0], 3), round(c[1], 3)) for c in curve[::10]], 'last', curve[-1])
    assert curve[-1][1] < curve[0][1] - 1.0                         # the model learns the four batches
    tr.close()


def test_load_state_dict_over_live_trainer_refreshes_r_attn_and_restarts_adam(monkeypatch):
    """fit -> load_state_dict(checkpoint) -> fit (resume): the first step after the reload must run with the NEW r_attn weights (their
    bf16 copy is training-only state) and with fresh Adam moments, exactly like a trainer created on the loaded weights."""
    cfg = small_config()
    bs, bptt = 2, 64
    om, pm, tr = build_pair(cfg, bs, bptt, 0.)
    g = torch.Generator().manual_seed(2)
    x, y = torch.randint(0, V, (bs, bptt), generator=g), torch.randint(0, V, (bs, bptt), generator=g)
    tr.reset()
    for _ in range(3):
        tr.step(x, y, lr=1e-2, mask_size=(1, 1))            # moves every weight, fills the Adam moments
    torch.manual_seed(123)
    om2 = txl.get_language_model(V, cfg, drop_mult=0.).train()      # a different "checkpoint"
    pm.load_state_dict(om2.state_dict())
    om2.reset(); tr.reset()
    set_oracle_mask(monkeypatch, (1, 1))
    opt = otrain.AdamTrueWD(otrain.unique_params(om2), eps=1e-3)
    ref = otrain.train_step(om2, x, y, opt, 1e-3, wd=0.01, clip=0.5)
    sd0 = {k: v.clone() for k, v in pm.state_dict().items()}
    tr.forward(x, y, None, mask_size=(1, 1)); tr.backward()
    got = tr.losses()
    assert abs(got['ce'] - ref['ce']) < 2e-2 * ref['ce'], (got, ref)
    compare_grads(tr, om2, tol=2e-2)                          # includes every r_attn.weight gradient
    tr.optimizer_step(1e-3, betas=(0.9, 0.99), eps=1e-3, wd=0.01, clip=0.5)
    sd_ref, sd_got = om2.state_dict(), pm.state_dict()
    for name, w in sd_got.items():                          # first Adam step of a FRESH optimizer (bias correction of step 1)
        d_got, d_ref = w - sd0[name], sd_ref[name].reshape(w.shape) - sd0[name]
        assert rel(d_got, d_ref) < 0.1, (name, rel(d_got, d_ref))
    tr.close()


def test_save_with_optimizer_state_and_reload_round_trip(tmp_path, golden_dir):
    """MusicLearner.save(file, config=...) -> createGenreContinuationModel(ckpt_path=file) -> predict (deep_music_genre.py:1784-1821,
    app_utils.py:68-75): weights, config and the Adam state come back; the reloaded learner generates the same greedy stream and its
    next training step equals the original learner's next step."""
    import os
    from deepmusicgeneration_b200.app_utils import createGenreContinuationModel
    from deepmusicgeneration_b200.codec import MusicDataBunch, MusicItem
    from deepmusicgeneration_b200.learner import music_model_learner
    cfg = small_config(mem_len=64)
    data = MusicDataBunch.empty('')
    learn = music_model_learner(data, config=dict(cfg), dtype='bf16', max_batch=4, max_seq=256, keep_hidden=False, seed=3)
    base = (torch.arange(4 * 64 * 8) % 29 + 12).view(4, -1)
    batches = [(base[:, i * 64:(i + 1) * 64], base[:, i * 64 + 1:(i + 1) * 64 + 1]) for i in range(7)]
    learn.fit_one_cycle(1, 3e-3, batches)
    path = str(tmp_path / 'ckpt.pth')
    learn.save(path, config=dict(cfg))
    state = torch.load(path, map_location='cpu', weights_only=False)
    assert set(state) == {'model', 'opt', 'config'} and state['opt'] is not None and state['config']['d_model'] == cfg['d_model']
    assert len(state['opt']['state']) == len(state['opt']['param_names']) > 10
    # reload through the app-level entry point (config comes from the file when the caller passes none: :1790-1791)
    learn2 = music_model_learner(data, config=None, pretrained_path=path, dtype='bf16', max_batch=4, max_seq=256, keep_hidden=False)
    for k, v in learn.model.state_dict().items():
        assert torch.equal(v, learn2.model.state_dict()[k]), k
    item = MusicItem.from_file(os.path.join(golden_dir, 'Undertale_-_Megalovania.mid'), data.vocab).trim_to_beat(8)
    a, _ = learn.predict(item, n_words=40, top_k=1, top_p=0.0, min_bars=100)
    b, _ = learn2.predict(item, n_words=40, top_k=1, top_p=0.0, min_bars=100)
    assert list(a.data) == list(b.data)
    # resume: one more step on both (same seed / step counter -> same dropout masks) moves the weights identically
    tr1 = learn.trainer(4, 64)
    learn2.load_opt_state(state['opt'], 4, 64)
    tr2 = learn2.trainer(4, 64)
    assert tr2.step_count == tr1.step_count
    for tr in (tr1, tr2):
        tr.reset(); tr.step(batches[0][0], batches[0][1], lr=1e-3, mask_size=(1, 1))
    sd1, sd2 = learn.model.state_dict(), learn2.model.state_dict()
    for k in sd1:
        assert (sd1[k] - sd2[k]).abs().max() < 1e-6, k
    assert createGenreContinuationModel.__defaults__[1].endswith('lakh_genre_model.pth')
