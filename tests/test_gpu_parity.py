"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical weights and inputs.

Tolerances: fp32 mode - logits within 5e-4 absolute of the oracle (different summation order only) and the greedy
token stream identical; bf16 mode - max relative logit error <= 2e-2 with rel = |a-b| / max(|b|, EPS), EPS = 1.0
(logits of the random-init model are O(1): std ~ 0.5, |max| ~ 2.5), top-1 agreement >= 99.9 %.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import bert as obert, codec as ocodec, sampling as osamp, txl

pytestmark = pytest.mark.gpu

V = 324
EPS_REL = 1.0
SMALL = dict(txl.default_config(), n_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256, mem_len=32,
             encode_position=False)
SMALL_M128 = dict(SMALL, mem_len=128)


def _pair(cfg, dtype, max_batch, max_seq, seed=0, keep_hidden=True, tame_unused=False, **kw):
    from deepmusicgeneration_b200.model import get_language_model
    torch.manual_seed(seed)
    om = txl.get_language_model(V, cfg).eval()
    if tame_unused:
        with torch.no_grad():
            om[1].decoder.bias[308:] = -50.
    pm = get_language_model(V, cfg, dtype=dtype, max_batch=max_batch, max_seq=max_seq, keep_hidden=keep_hidden, init=False, **kw)
    pm.load_state_dict(om.state_dict())
    return om, pm


def _rel(a, b):
    return ((a - b).abs() / b.abs().clamp_min(EPS_REL)).max().item()


def test_txl_f32_forward_segments_and_mems():
    om, pm = _pair(SMALL, 'f32', 3, 64)
    g = torch.Generator().manual_seed(1234)
    om.reset(); pm.reset()
    for T in (50, 1, 1, 1, 7, 33, 1):
        x = torch.randint(0, V, (3, T), generator=g)
        with torch.no_grad():
            ol, ohid, oout = om(x)
        pl, phid, pout = pm(x.cuda())
        assert (pl.cpu() - ol).abs().max() < 5e-4, T
        assert (pout[0].cpu() - oout[0]).abs().max() < 5e-4, T
        assert len(phid) == len(ohid)
        for a, b in zip(phid, ohid):
            assert a.shape == b.shape, (a.shape, b.shape)
            assert (a.cpu() - b).abs().max() < 5e-4


def test_txl_f32_encode_position_and_window_mask():
    cfg = dict(SMALL, encode_position=True, mask_steps=4)
    om, pm = _pair(cfg, 'f32', 2, 64)
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, V, (2, 40), generator=g)
    pos = torch.cumsum(torch.randint(0, 40, (2, 40), generator=g), 1)
    om.reset(); pm.reset()
    with torch.no_grad():
        ol = om({'x': x, 'pos': pos.clone()})[0]
    pl = pm({'x': x.cuda(), 'pos': pos.cuda()})[0]
    assert (pl.cpu() - ol).abs().max() < 5e-4
    # training-time window mask (deep_music_genre.py:1577-1590), forced to (win=3, k=0) on both sides, with memory
    x2 = torch.randint(0, V, (2, 20), generator=g)
    pos2 = pos[:, -1:] + torch.cumsum(torch.randint(0, 9, (2, 20), generator=g), 1)

    class _Rng:
        def rand(self): return 0.0
        def randint(self, lo, hi): return 2
    orig = txl.rand_window_mask
    txl.rand_window_mask = lambda x_len, m_len, device, max_size=None, p=0.2, is_eval=False: orig(x_len, m_len, device, max_size=max_size, p=p, is_eval=False, rng=_Rng())
    try:
        with torch.no_grad():
            ol2 = om({'x': x2, 'pos': pos2.clone()})[0]
    finally:
        txl.rand_window_mask = orig
    pl2 = pm({'x': x2.cuda(), 'pos': pos2.cuda()}, mask_size=(3, 0))[0]
    assert (pl2.cpu() - ol2).abs().max() < 5e-4


def test_txl_f32_baseline_config():
    om, pm = _pair(txl.baseline_config(), 'f32', 1, 640, keep_hidden=False)
    g = torch.Generator().manual_seed(5)
    om.reset(); pm.reset()
    for T in (96, 1, 1):
        x = torch.randint(0, V, (1, T), generator=g)
        with torch.no_grad():
            ol = om(x)[0]
        pl = pm[0].forward(x.cuda(), logits_mode=1)[0]
        assert (pl.cpu() - ol).abs().max() < 1e-3, T
        assert (pl.cpu().argmax(-1) == ol.argmax(-1)).all()


def test_txl_bf16_logits_and_top1():
    B, T = 8, 512
    om, pm = _pair(txl.baseline_config(), 'bf16', B, 512, keep_hidden=False)
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, V, (B, T), generator=g)
    om.reset(); pm.reset()
    with torch.no_grad():
        ol = om(x)[0]
    pl = pm[0].forward(x.cuda(), logits_mode=1)[0].cpu()
    rel = _rel(pl, ol)
    top1 = (pl.argmax(-1) == ol.argmax(-1)).float().mean().item()
    print(f'bf16 prefill: max rel err {rel:.4e} (eps {EPS_REL}), max abs {(pl - ol).abs().max():.4e}, logits absmax {ol.abs().max():.3f} '
          f'std {ol.std():.3f}, top-1 agreement {top1:.5f} over {B * T} positions')
    assert rel <= 2e-2
    # top-1: positions whose oracle top-2 margin is below the bf16 noise floor are ties, not disagreements
    srt = ol.sort(-1, descending=True)[0]
    clear = (srt[..., 0] - srt[..., 1]) > 2 * (pl - ol).abs().max()
    top1_clear = (pl.argmax(-1) == ol.argmax(-1))[clear].float().mean().item()
    print(f'top-1 on {int(clear.sum())} clear-margin positions: {top1_clear:.5f}')
    assert top1 >= 0.99 and top1_clear >= 0.999
    # decode continues from the bf16 ring: 8 single-token steps against the oracle
    for s in range(8):
        xs = torch.randint(0, V, (B, 1), generator=g)
        with torch.no_grad():
            o1 = om(xs)[0]
        p1 = pm[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        assert _rel(p1, o1) <= 2e-2, s


def test_decode_lanes_match_single_lane():
    "DMG_DECODE_LANES=2: the one-token step issued as two groups of streams on parallel streams (opt-in) gives the same logits"
    cfg = dict(txl.baseline_config(), n_layers=2)
    om, pa = _pair(cfg, 'bf16', 160, 64, keep_hidden=False)
    _, pb = _pair(cfg, 'bf16', 160, 64, keep_hidden=False)
    g = torch.Generator().manual_seed(21)
    x0 = torch.randint(0, V, (160, 30), generator=g)
    for pm in (pa, pb):
        pm.reset(); pm[0].forward(x0.cuda(), logits_mode=2)
    worst = 0.
    for s in range(5):
        xs = torch.randint(0, V, (160, 1), generator=g)
        os.environ.pop('DMG_DECODE_LANES', None)
        la = pa[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        try:
            os.environ['DMG_DECODE_LANES'] = '2'
            lb = pb[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        finally:
            os.environ.pop('DMG_DECODE_LANES', None)
        worst = max(worst, (la - lb).abs().max().item())
    assert worst < 1e-3, worst


def test_scaled_model_c5_shapes_prefill_and_decode():
    "BASELINE.json configs[4]: d_model 1024, 16 heads x 64, d_inner 4096, mem_len 1024 (2 of the 24 layers): prefill + ring decode"
    cfg = dict(txl.baseline_config(), n_layers=2, d_model=1024, n_heads=16, d_head=64, d_inner=4096, mem_len=1024, ctx_len=1024)
    om, pm = _pair(cfg, 'bf16', 3, 1024, keep_hidden=False)
    g = torch.Generator().manual_seed(5)
    x0 = torch.randint(0, V, (3, 1000), generator=g)
    om.reset(); pm.reset()
    with torch.no_grad(): ol = om(x0)[0][:, -1]
    pl = pm[0].forward(x0.cuda(), logits_mode=2)[0].cpu()
    assert _rel(pl, ol) <= 2e-2
    worst = 0.
    for s in range(40):                                        # crosses the point where the 1024-slot ring wraps
        xs = torch.randint(0, V, (3, 1), generator=g)
        with torch.no_grad(): lo = om(xs)[0][:, -1]
        lp = pm[0].forward(xs.cuda(), logits_mode=1)[0].cpu()[:, -1]
        worst = max(worst, _rel(lp, lo))
    print(f'C5 shapes: decode max rel err {worst:.3e}')
    assert worst <= 2e-2


def test_fused_gemm_layernorm_cluster_kernel_matches_unfused_pair():
    "DMG_GEMM_LN=1: out-projection / FFN-down GEMM + residual + LayerNorm in one 8-CTA-cluster kernel (gemm_ln.cu, opt-in)"
    cfg = dict(txl.baseline_config(), n_layers=3)
    om, pa = _pair(cfg, 'bf16', 130, 64, keep_hidden=False)       # 130 streams: two row tiles, the second one ragged
    _, pb = _pair(cfg, 'bf16', 130, 64, keep_hidden=False)
    g = torch.Generator().manual_seed(11)
    x0 = torch.randint(0, V, (130, 40), generator=g)
    om.reset()
    with torch.no_grad(): om(x0)
    for pm in (pa, pb):
        pm.reset(); pm[0].forward(x0.cuda(), logits_mode=2)
    worst = worst_o = 0.
    for s in range(6):
        xs = torch.randint(0, V, (130, 1), generator=g)
        os.environ.pop('DMG_GEMM_LN', None)
        la = pa[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        try:
            os.environ['DMG_GEMM_LN'] = '1'
            lb = pb[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        finally:
            os.environ.pop('DMG_GEMM_LN', None)
        with torch.no_grad(): lo = om(xs)[0]
        worst = max(worst, (la - lb).abs().max().item())
        worst_o = max(worst_o, _rel(lb, lo))
    print(f'fused GEMM+LN vs unfused: max abs {worst:.3e}; vs oracle max rel {worst_o:.3e}')
    assert worst < 1e-2 and worst_o <= 2e-2


@pytest.mark.parametrize('M', [128, 64, 512])
def test_decode_kernels_equal_general_kernel_and_oracle(M):
    """x_len==1 fast kernels (v2: persistent, TMA-2D tiles, resident Rd, mma.sync; v1: bulk-copy ring + FFMA) vs the
    general kernel vs the fp32 oracle, from a partly filled memory across several ring wrap-arounds."""
    cfg = dict(SMALL, mem_len=M)
    om, p2 = _pair(cfg, 'bf16', 5, 128, keep_hidden=False)
    _, p1 = _pair(cfg, 'bf16', 5, 128, keep_hidden=False)
    _, pg = _pair(cfg, 'bf16', 5, 128, keep_hidden=False)
    g = torch.Generator().manual_seed(3)
    T0 = min(100, M - 28)
    x0 = torch.randint(0, V, (5, T0), generator=g)              # memory only partly filled: masked ring slots
    om.reset()
    with torch.no_grad(): om(x0)
    for pm in (p2, p1, pg):
        pm.reset(); pm[0].forward(x0.cuda(), logits_mode=2)
    w2g = w1g = w2o = 0.
    n_steps = 2 * M + 44 if M <= 128 else M + 40
    for s in range(n_steps):
        xs = torch.randint(0, V, (5, 1), generator=g)
        for k in ('DMG_NO_DECODE_KERNEL', 'DMG_DECODE_V1'): os.environ.pop(k, None)
        l2 = p2[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        try:
            os.environ['DMG_DECODE_V1'] = '1'
            l1 = p1[0].forward(xs.cuda(), logits_mode=1)[0].cpu() if M % 128 == 0 else None
            os.environ.pop('DMG_DECODE_V1')
            os.environ['DMG_NO_DECODE_KERNEL'] = '1'
            lg = pg[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        finally:
            for k in ('DMG_NO_DECODE_KERNEL', 'DMG_DECODE_V1'): os.environ.pop(k, None)
        with torch.no_grad(): lo = om(xs)[0]
        w2g = max(w2g, (l2 - lg).abs().max().item())
        if l1 is not None: w1g = max(w1g, (l1 - lg).abs().max().item())
        w2o = max(w2o, _rel(l2, lo))
    print(f'M={M}: decode v2 vs general max abs {w2g:.3e}; v1 vs general {w1g:.3e}; v2 vs oracle max rel {w2o:.3e}')
    assert w2g < 2e-2 and w1g < 2e-2 and w2o <= 2e-2


def test_greedy_token_stream_f32_bit_exact(golden_dir):
    "fp32 mode: the greedy stream of MusicLearner.predict equals the oracle's reference loop token for token."
    from deepmusicgeneration_b200.codec import MusicDataBunch, MusicItem
    from deepmusicgeneration_b200.learner import MusicLearner
    for cfg, n_words, beat in ((SMALL_M128, 300, 8), (txl.baseline_config(), 40, 16)):
        om, pm = _pair(cfg, 'f32', 1, 1024, keep_hidden=False, tame_unused=True)
        data = MusicDataBunch.empty('')
        item = MusicItem.from_file(os.path.join(golden_dir, 'Undertale_-_Megalovania.mid'), data.vocab).trim_to_beat(beat)
        item.data[0] = data.vocab.stoi['xxpop']
        ov = ocodec.MusicVocab.create()
        ref = osamp.predict(om, ov, item.data, item.position, n_words=n_words, temperatures=(1.3, 1.1, 0.9), min_bars=12,
                            top_k=1, top_p=0.0)
        learn = MusicLearner(data, pm)
        pred, full = learn.predict(item, n_words=n_words, temperatures=(1.3, 1.1, 0.9), min_bars=12, top_k=1, top_p=0.0)
        assert list(pred.data) == ref, (cfg['n_layers'], len(pred.data), len(ref))
        assert list(full.data) == list(item.data) + ref


def test_generate_graph_replay_equals_eager():
    from deepmusicgeneration_b200.codec import MusicDataBunch
    from deepmusicgeneration_b200.learner import MusicLearner
    om, pm = _pair(SMALL_M128, 'bf16', 6, 128, keep_hidden=False, tame_unused=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randint(12, 140, (6, 60), generator=g); x[:, -1] = 301     # seeds end on an instrument token
    learn = MusicLearner(MusicDataBunch.empty(''), pm)
    a = learn.generate_batch(x, n_words=150, top_k=1, top_p=0.0, min_bars=10 ** 6).cpu()
    os.environ['DMG_NO_GRAPH'] = '1'
    try:
        b = learn.generate_batch(x, n_words=150, top_k=1, top_p=0.0, min_bars=10 ** 6).cpu()
    finally:
        os.environ.pop('DMG_NO_GRAPH', None)
    assert torch.equal(a, b)
    assert (a >= 0).all()


def test_select_hidden_permutes_streams():
    om, pm = _pair(SMALL, 'f32', 4, 64)
    g = torch.Generator().manual_seed(2)
    x = torch.randint(0, V, (3, 20), generator=g)
    om.reset(); pm.reset()
    with torch.no_grad(): om(x)
    pm(x.cuda())
    idx = torch.tensor([2, 0, 0, 1])
    om[0].select_hidden(idx); pm[0].select_hidden(idx)
    x2 = torch.randint(0, V, (4, 3), generator=g)
    with torch.no_grad(): ol = om(x2)[0]
    pl = pm(x2.cuda())[0].cpu()
    assert (pl - ol).abs().max() < 5e-4


def _bert_pair(cfg, dtype, max_batch, max_seq, seed=0):
    from deepmusicgeneration_b200.model import get_multitask_model
    torch.manual_seed(seed)
    om = obert.get_multitask_model(V, cfg, pad_idx=1).eval()
    pm = get_multitask_model(V, cfg, pad_idx=1, dtype=dtype, max_batch=max_batch, max_seq=max_seq, init=False)
    pm.load_state_dict(om.state_dict())
    return om, pm


@pytest.mark.parametrize('T', [1, 2, 31, 70, 130])
def test_bert_encoder_f32_wraparound_live(T):
    cfg = dict(obert.multitask_config(), enc_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256)
    om, pm = _bert_pair(cfg, 'f32', 2, 256)
    g = torch.Generator().manual_seed(T)
    x = torch.randint(0, V, (2, T), generator=g)
    pos = torch.cumsum(torch.randint(0, 9, (2, T), generator=g), 1)
    with torch.no_grad():
        ol = om({'msk': {'x': x, 'pos': pos.clone()}})['msk']
    pl = pm({'msk': {'x': x.cuda(), 'pos': pos.cuda()}})['msk'].cpu()
    assert pl.shape == ol.shape
    assert (pl - ol).abs().max() < 5e-4


def test_bert_encoder_bf16_app_config():
    om, pm = _bert_pair(obert.multitask_config(), 'bf16', 2, 512)
    g = torch.Generator().manual_seed(9)
    x = torch.randint(0, V, (2, 400), generator=g)
    pos = torch.cumsum(torch.randint(0, 9, (2, 400), generator=g), 1)
    with torch.no_grad():
        ol = om({'msk': {'x': x, 'pos': pos.clone()}})['msk']
    pl = pm({'msk': {'x': x.cuda(), 'pos': pos.cuda()}})['msk'].cpu()
    rel = _rel(pl, ol)
    top1 = (pl.argmax(-1) == ol.argmax(-1)).float().mean().item()
    print(f'bert bf16: max rel err {rel:.3e}, top-1 {top1:.4f}')
    assert rel <= 2e-2 and top1 >= 0.99


@pytest.mark.parametrize('T', [2, 31, 64, 70, 130, 257, 320])
def test_bert_flash_attention_bf16_ragged_lengths(T):
    "attention_flash.cu, BERT mode (all three _line_shift lines live, ragged last tile) against the oracle and the general kernel"
    cfg = dict(obert.multitask_config(), enc_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256)
    om, pm = _bert_pair(cfg, 'bf16', 3, 320)
    g = torch.Generator().manual_seed(T)
    x = torch.randint(0, V, (3, T), generator=g)
    pos = torch.cumsum(torch.randint(0, 9, (3, T), generator=g), 1)
    with torch.no_grad():
        ol = om({'msk': {'x': x, 'pos': pos.clone()}})['msk']
    os.environ.pop('DMG_NO_FLASH', None)
    try:
        os.environ['DMG_BERT_ATTN_MMA_SYNC'] = '1'       # sequences of 128 tokens and more default to attention_bert_tc.cu
        pl = pm({'msk': {'x': x.cuda(), 'pos': pos.cuda()}})['msk'].cpu()
        os.environ['DMG_NO_FLASH'] = '1'
        pg = pm({'msk': {'x': x.cuda(), 'pos': pos.cuda()}})['msk'].cpu()
    finally:
        os.environ.pop('DMG_NO_FLASH', None)
        os.environ.pop('DMG_BERT_ATTN_MMA_SYNC', None)
    print(f'T={T}: flash vs oracle {_rel(pl, ol):.3e}, general vs oracle {_rel(pg, ol):.3e}, flash vs general {(pl - pg).abs().max():.3e}')
    assert _rel(pl, ol) <= 2e-2 and (pl - pg).abs().max() < 3e-2


@pytest.mark.parametrize('variant', ['default', 'DMG_BERT_TC_FP32_STRIP', 'DMG_BERT_TC16'])
@pytest.mark.parametrize('T', [128, 129, 191, 256, 257, 320, 384, 1000, 1024])
def test_bert_tcgen05_attention_bf16(T, variant):
    """attention_bert_tc.cu (tcgen05 / TMEM / TMA; sequences of 128 tokens and more): all three _line_shift lines, the zero pad at
    j = i + 1 (also across a tile boundary), the wrapped line-3 distances and a ragged last tile (masked keys, one or both key
    halves), against the oracle and the FFMA general kernel.  Variants: the default (fp16 strip line), the fp32 strip line
    (DMG_BERT_TC_FP32_STRIP) and the sixteen-softmax-warp kernel (attention_bert_tc16.cu, DMG_BERT_TC16)"""
    cfg = dict(obert.multitask_config(), enc_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256)
    om, pm = _bert_pair(cfg, 'bf16', 3, 1024)
    g = torch.Generator().manual_seed(T)
    x = torch.randint(0, V, (3, T), generator=g)
    pos = torch.cumsum(torch.randint(0, 9, (3, T), generator=g), 1)
    with torch.no_grad():
        ol = om({'msk': {'x': x, 'pos': pos.clone()}})['msk']
    os.environ.pop('DMG_NO_FLASH', None)
    try:
        if variant != 'default':
            os.environ[variant] = '1'
        pl = pm({'msk': {'x': x.cuda(), 'pos': pos.cuda()}})['msk'].cpu()
        os.environ['DMG_NO_FLASH'] = '1'
        pg = pm({'msk': {'x': x.cuda(), 'pos': pos.cuda()}})['msk'].cpu()
    finally:
        os.environ.pop('DMG_NO_FLASH', None)
        os.environ.pop(variant, None)
    print(f'T={T}: tcgen05 vs oracle {_rel(pl, ol):.3e}, general vs oracle {_rel(pg, ol):.3e}, tcgen05 vs general {(pl - pg).abs().max():.3e}')
    assert _rel(pl, ol) <= 2e-2 and (pl - pg).abs().max() < 3e-2


@pytest.mark.parametrize('T', [2, 63, 64, 100, 300, 320])
def test_txl_flash_prefill_bf16_ragged_lengths(T):
    "attention_flash.cu, causal mode: prefill of a ragged seed after reset(), then ring decode continues from its K/V"
    cfg = dict(SMALL, mem_len=128)
    om, pm = _pair(cfg, 'bf16', 3, 320, keep_hidden=False)
    _, pg = _pair(cfg, 'bf16', 3, 320, keep_hidden=False)
    g = torch.Generator().manual_seed(100 + T)
    x0 = torch.randint(0, V, (3, T), generator=g)
    om.reset(); pm.reset(); pg.reset()
    with torch.no_grad(): ol = om(x0)[0]
    os.environ.pop('DMG_NO_FLASH', None)
    pl = pm[0].forward(x0.cuda(), logits_mode=1)[0].cpu()
    try:
        os.environ['DMG_NO_FLASH'] = '1'
        gl = pg[0].forward(x0.cuda(), logits_mode=1)[0].cpu()
    finally:
        os.environ.pop('DMG_NO_FLASH', None)
    assert _rel(pl, ol) <= 2e-2 and (pl - gl).abs().max() < 3e-2
    for s in range(5):
        xs = torch.randint(0, V, (3, 1), generator=g)
        with torch.no_grad(): lo = om(xs)[0]
        lp = pm[0].forward(xs.cuda(), logits_mode=1)[0].cpu()
        assert _rel(lp, lo) <= 2e-2, s


def test_predict_mask_greedy_matches_oracle(golden_dir):
    from deepmusicgeneration_b200.codec import MusicDataBunch, MusicItem
    from deepmusicgeneration_b200.learner import MultitaskLearner
    cfg = dict(obert.multitask_config(), enc_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256)
    om, pm = _bert_pair(cfg, 'f32', 1, 256)
    data = MusicDataBunch.empty('')
    item = MusicItem.from_file(os.path.join(golden_dir, 'uploadedMidi.mid'), data.vocab).trim_to_beat(8)
    notes = [i for i, t in enumerate(item.data) if data.vocab.note_range[0] <= t < data.vocab.note_range[1]]
    item.data[notes[::2]] = data.vocab.mask_idx
    ref = osamp.predict_mask(om, ocodec.MusicVocab.create(), item.data, item.position, temperatures=(1.1, 0.9), top_k=1, top_p=0.0)
    out = MultitaskLearner(data, pm).predict_mask(item, temperatures=(1.1, 0.9), top_k=1, top_p=0.0)
    assert list(out.data) == list(ref)
