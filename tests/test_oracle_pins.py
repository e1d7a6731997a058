"""The oracle against every known answer the reference holds for the hot path."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bert, codec, sampling, txl


@pytest.fixture(scope='module')
def pins(golden_dir):
    return json.load(open(os.path.join(golden_dir, 'notebook_pins.json')))


def test_vocab_size_and_layout(pins):
    v = codec.MusicVocab.create()
    assert len(v) == pins['vocab_size'] == 324
    assert v.stoi['xxni'] == pins['init_prev_idx'] == 10          # notebook cell 81 prints "Init prev_idx =  10"
    assert v.note_range == (12, 140) and v.dur_range == (140, 301) and v.ins_range == (301, 308)
    assert v.itos[-6:] == [f'dummy{i}' for i in range(6)]


def test_parameter_count_pins_architecture(pins):
    torch.manual_seed(0)
    m = txl.get_language_model(324, txl.btp_phase1_config())
    assert txl.count_parameters(m) == pins['btp_phase1_param_count'] == 41107268
    m16 = txl.get_language_model(324, txl.baseline_config())
    assert txl.count_parameters(m16) == 54766916                  # SURVEY.md App. B, BASELINE C1-C3 model


def test_state_dict_keys_follow_fastai():
    m = txl.get_language_model(324, dict(txl.default_config(), n_layers=1))
    keys = set(m.state_dict().keys())
    for k in ['0.encoder.weight', '0.u', '0.v', '0.pos_enc.freq', '0.layers.0.mhra.attention.weight',
              '0.layers.0.mhra.out.weight', '0.layers.0.mhra.r_attn.weight', '0.layers.0.mhra.ln.weight',
              '0.layers.0.ff.layers.0.weight', '0.layers.0.ff.layers.3.bias', '0.layers.0.ff.layers.6.weight',
              '1.decoder.weight', '1.decoder.bias', '0.beat_enc.beat_enc.weight', '0.beat_enc.bar_enc.weight']:
        assert k in keys, k
    assert m[1].decoder.weight is m[0].encoder.weight             # tied head


def test_megalovania_golden_tokens(golden_dir):
    v = codec.MusicVocab.create()
    gold = open(os.path.join(golden_dir, 'megalovania_seed64.txt')).read().split()
    idx = codec.seed_from_midi(os.path.join(golden_dir, 'Undertale_-_Megalovania.mid'), v, cutoff_beat=64,
                               genre_token='xxelec')
    assert len(idx) == 623
    assert v.textify(idx).split(' ') == gold


def test_line_shift_is_index_arithmetic():
    "SURVEY.md App. A.4: the three regions of _line_shift."
    for T, M in [(6, 0), (5, 4), (1, 7), (9, 3)]:
        S = M + T
        raw = torch.arange(T * S, dtype=torch.float32).view(1, 1, T, S) + 1
        out = txl._line_shift(raw)[0, 0]
        for i in range(T):
            for j in range(S):
                if j <= M + i:
                    exp = raw[0, 0, i, j + T - 1 - i]
                elif j == M + i + 1:
                    exp = 0.
                else:
                    exp = raw[0, 0, i + 1, j - i - M - 2]
                assert out[i, j] == exp, (T, M, i, j)


def test_incremental_equals_full_context():
    torch.manual_seed(1)
    cfg = dict(txl.default_config(), n_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256, mem_len=64,
               encode_position=False)
    m = txl.get_language_model(324, cfg).eval()
    x = torch.randint(0, 324, (2, 40))
    with torch.no_grad():
        m.reset(); full = m(x)[0]
        m.reset(); inc = torch.cat([m(x[:, :25])[0], m(x[:, 25:26])[0], m(x[:, 26:])[0]], 1)
    assert (full - inc).abs().max() < 1e-4


def test_eval_mask_is_causal_with_memory_visible():
    mask = txl.rand_window_mask(5, 3, 'cpu', max_size=4, is_eval=True)[0, 0]
    exp = torch.triu(torch.ones(5, 8), diagonal=4).bool()
    assert torch.equal(mask, exp)
    wm = txl.window_mask(6, 'cpu', m_len=2, size=(2, 0))[0, 0]
    for i in range(6):
        for jx in range(6):
            assert bool(wm[i, 2 + jx]) == (not (jx == 0 or jx // 2 < i // 2))


def test_top_k_top_p_reference_semantics():
    logits = torch.tensor([2.0, 1.0, 0.5, 0.0, -1.0])
    out = sampling.top_k_top_p(logits, top_k=3, top_p=0.0)
    assert torch.isinf(out[3:]).all() and not torch.isinf(out[:3]).any()
    out = sampling.top_k_top_p(logits, top_k=0, top_p=0.5)          # cum = [.57, ...] -> keep first only... plus shift
    assert not torch.isinf(out[0]) and torch.isinf(out[2:]).all()
    out = sampling.top_k_top_p(logits, top_k=1, top_p=0.0)
    assert (~torch.isinf(out)).sum() == 1 and out.argmax() == 0


def test_grammar_filter_classes():
    v = codec.MusicVocab.create()
    free = lambda prev, **kw: set(torch.isfinite(sampling.filter_invalid_indexes(torch.zeros(324), prev, v, **kw)).nonzero().view(-1).tolist())
    unused = set(range(308, 324))                                    # mt*, dummy*: the reference never filters them
    assert free(v.stoi['d4']) == set(range(*v.ins_range)) | unused
    assert free(v.stoi['d4'], last_xxsep=True) == {v.ni_idx} | unused
    assert free(v.stoi['i0']) == set(range(*v.note_range)) | {v.sep_idx} | unused
    assert free(v.stoi['n60']) == set(range(*v.dur_range)) | unused
    assert free(v.sep_idx) == set(range(*v.dur_range)) | unused


def test_oracle_predict_greedy_is_deterministic_and_grammatical(golden_dir):
    torch.manual_seed(0)
    v = codec.MusicVocab.create()
    cfg = dict(txl.default_config(), n_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256, mem_len=64,
               encode_position=False)
    m = txl.get_language_model(len(v), cfg).eval()
    # push mt*/dummy* logits down so the random-init model follows the n-d-i grammar like a trained one
    with torch.no_grad(): m[1].decoder.bias[308:] = -50.
    seed = codec.seed_from_midi(os.path.join(golden_dir, 'Undertale_-_Megalovania.mid'), v, cutoff_beat=8, genre_token='xxpop')
    pos = codec.position_enc(seed, v)
    a = sampling.predict(m, v, seed, pos, n_words=24, top_k=1, top_p=0.0, min_bars=100)
    b = sampling.predict(m, v, seed, pos, n_words=24, top_k=1, top_p=0.0, min_bars=100)
    assert a == b and len(a) > 0
    prev = int(seed[-1])
    for t in a:
        if v.is_duration(prev): assert v.is_ins(t)
        elif v.is_ins(prev): assert v.note_range[0] <= t < v.note_range[1] or t == v.sep_idx
        else: assert v.is_duration(t)
        prev = t


def test_bert_oracle_shapes_and_wraparound_live():
    torch.manual_seed(0)
    cfg = dict(bert.multitask_config(), enc_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256)
    m = bert.get_multitask_model(324, cfg, pad_idx=1).eval()
    x = torch.randint(0, 324, (2, 12)); pos = torch.cumsum(torch.randint(0, 4, (2, 12)), 1)
    with torch.no_grad():
        out = m({'msk': {'x': x, 'pos': pos}})['msk']
    assert out.shape == (2, 12, 324)
    keys = set(m.state_dict().keys())
    assert 'encoder.layers.0.mha1.q_wgt.bias' in keys and 'head.decoder.bias' in keys and 'encoder.embed.bar_enc.weight' in keys
    # no attention mask in the encoder: changing a LATER token changes EARLIER positions (bidirectional)
    x2 = x.clone(); x2[:, -1] = (x2[:, -1] + 1) % 324
    with torch.no_grad():
        out2 = m({'msk': {'x': x2, 'pos': pos}})['msk']
    assert (out2[:, 0] - out[:, 0]).abs().max() > 0
