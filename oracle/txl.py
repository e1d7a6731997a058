"""Oracle: fastai==1.0.61 Transformer-XL language model as the reference uses it.  TEST INFRASTRUCTURE.

Plain PyTorch fp32 on the CPU.  The arithmetic lives in an UN-VENDORED dependency of the reference
(fastai==1.0.61, ``fastai/text/models/transformer.py``, ``fastai/text/models/awd_lstm.py``,
``fastai/text/learner.py``; pinned by ``notebooks/Transformer_Genre_Evaluation.ipynb:162``) and is
restated here from its published source; the in-repo parts follow
``deep_music_genre.py:1577-1665`` (masks, ``MusicTransformerXL.forward``, ``BeatPositionEncoder``).
The attention body is cross-checked against the reference's in-repo twin
``MemMultiHeadRelativeAttentionKV._apply_attention`` (``deep_music_remix.py:2078-2104``).

Module/attribute names mirror fastai so that ``state_dict()`` keys equal the reference checkpoint
keys (``0.encoder.weight``, ``0.layers.3.mhra.attention.weight``, ``0.layers.3.ff.layers.0.weight``,
``1.decoder.bias`` ...).

Parity status: architecture pinned by the 41,107,268-parameter known answer; logits unpinned (see
``oracle/__init__.py``).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# configs (fastai tfmerXL_lm_config + the reference's overrides, app_utils.py:13-63)
# ---------------------------------------------------------------------------------------------
def tfmerXL_lm_config():
    "fastai.text.models.transformer.tfmerXL_lm_config (act is overridden to GeLU by every reference config)."
    return dict(ctx_len=150, n_layers=12, n_heads=10, d_model=410, d_head=41, d_inner=2100, resid_p=0.1,
                attn_p=0.1, ff_p=0.1, embed_p=0.1, output_p=0.1, bias=False, scale=True, act='relu',
                double_drop=True, tie_weights=True, out_bias=True, mem_len=150, mask=True)


def default_config():
    "app_utils.py:13-24"
    c = tfmerXL_lm_config()
    c.update(act='gelu', mem_len=512, d_model=512, d_inner=2048, n_layers=6, n_heads=8, d_head=64)
    return c


def music_config():
    "app_utils.py:26-38"
    c = default_config()
    c['ctx_len'] = 512
    return c


def btp_phase1_config():
    "app_utils.py:40-53 (the genre app model; 41,107,268 parameters with V=324)"
    c = default_config()
    c.update(ctx_len=512, d_inner=3072, n_heads=12, d_head=64, n_layers=8, transpose_range=(0, 12),
             mask_steps=4, encode_position=False)
    return c


def baseline_config():
    "BASELINE.json C1-C3 'musicautobot default': d 512, 16 layers, 8 heads, mem 512."
    c = default_config()
    c.update(n_layers=16, ctx_len=512, encode_position=False, mask_steps=1)
    return c


# ---------------------------------------------------------------------------------------------
# fastai building blocks
# ---------------------------------------------------------------------------------------------
class GeLU(nn.Module):
    "fastai GeLU: the tanh form, NOT erf."
    def forward(self, x):
        return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


class PositionalEncoding(nn.Module):
    "fastai PositionalEncoding: cat(sin, cos) halves, freq_k = 1/10000^(2k/d)."
    def __init__(self, d):
        super().__init__()
        self.register_buffer('freq', 1 / (10000 ** (torch.arange(0., d, 2.) / d)))

    def forward(self, pos):
        inp = torch.outer(pos, self.freq)
        return torch.cat([inp.sin(), inp.cos()], dim=-1)


def _line_shift(x, mask=False):
    "fastai _line_shift: shift line i of `x` by p-i elements to the left (pad, view, drop row 0, view back)."
    bs, nh, n, p = x.size()
    x_pad = torch.cat([x.new_zeros(bs, nh, n, 1), x], dim=3)
    x_shift = x_pad.view(bs, nh, p + 1, n)[:, :, 1:].reshape(bs, nh, n, p)
    if mask:
        x_shift = x_shift * torch.tril(x.new_ones(n, p), p - n)[None, None]
    return x_shift


class MergeLayer(nn.Module):
    def forward(self, x, orig):
        return x + orig


class SequentialEx(nn.Module):
    "fastai SequentialEx: like nn.Sequential but MergeLayer sees the block input."
    def __init__(self, *layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        res = x
        for l in self.layers:
            res = l(res, x) if isinstance(l, MergeLayer) else l(res)
        return res


def feed_forward(d_model, d_ff, ff_p=0., act='relu', double_drop=True):
    "fastai feed_forward: Linear, act, [Dropout], Linear, Dropout, Merge, LayerNorm (indices 0..6)."
    layers = [nn.Linear(d_model, d_ff), GeLU() if act == 'gelu' else nn.ReLU()]
    if double_drop:
        layers.append(nn.Dropout(ff_p))
    return SequentialEx(*layers, nn.Linear(d_ff, d_model), nn.Dropout(ff_p), MergeLayer(), nn.LayerNorm(d_model))


class MultiHeadRelativeAttention(nn.Module):
    "fastai MultiHeadRelativeAttention (subclass of MultiHeadAttention; forward = ln(x + drop(out(attn))))."
    def __init__(self, n_heads, d_model, d_head=None, resid_p=0., attn_p=0., bias=True, scale=True):
        super().__init__()
        d_head = d_head if d_head is not None else d_model // n_heads
        self.n_heads, self.d_head, self.scale = n_heads, d_head, scale
        self.attention = nn.Linear(d_model, 3 * n_heads * d_head, bias=bias)
        self.out = nn.Linear(n_heads * d_head, d_model, bias=bias)
        self.drop_att, self.drop_res = nn.Dropout(attn_p), nn.Dropout(resid_p)
        self.ln = nn.LayerNorm(d_model)
        self.r_attn = nn.Linear(d_model, n_heads * d_head, bias=bias)

    def forward(self, x, mask=None, **kwargs):
        return self.ln(x + self.drop_res(self.out(self._apply_attention(x, mask=mask, **kwargs))))

    def _apply_attention(self, x, r=None, u=None, v=None, mask=None, mem=None):
        bs, x_len, seq_len = x.size(0), x.size(1), r.size(0)
        # legacy torch.cat([empty_1d, x3d], 1) == x  
        context = x if (mem is None or mem.dim() == 1) else torch.cat([mem, x], dim=1)
        wq, wk, wv = torch.chunk(self.attention(context), 3, dim=-1)
        wq = wq[:, -x_len:]
        wq, wk, wv = map(lambda t: t.view(bs, t.size(1), self.n_heads, self.d_head), (wq, wk, wv))
        wq, wk, wv = wq.permute(0, 2, 1, 3), wk.permute(0, 2, 3, 1), wv.permute(0, 2, 1, 3)
        wkr = self.r_attn(r)
        wkr = wkr.view(seq_len, self.n_heads, self.d_head)
        wkr = wkr.permute(1, 2, 0)
        AC = torch.matmul(wq + u, wk)
        BD = _line_shift(torch.matmul(wq + v, wkr))
        attn_score = AC + BD
        if self.scale:
            attn_score = attn_score.mul_(1 / (self.d_head ** 0.5))
        if mask is not None:
            attn_score = attn_score.float().masked_fill(mask, -float('inf')).type_as(attn_score)
        attn_prob = self.drop_att(F.softmax(attn_score, dim=-1))
        attn_vec = torch.matmul(attn_prob, wv)
        return attn_vec.permute(0, 2, 1, 3).contiguous().view(bs, x_len, -1)


class DecoderLayer(nn.Module):
    def __init__(self, n_heads, d_model, d_head, d_inner, resid_p=0., attn_p=0., ff_p=0., bias=True, scale=True,
                 act='relu', double_drop=True):
        super().__init__()
        self.mhra = MultiHeadRelativeAttention(n_heads, d_model, d_head, resid_p=resid_p, attn_p=attn_p,
                                               bias=bias, scale=scale)
        self.ff = feed_forward(d_model, d_inner, ff_p=ff_p, act=act, double_drop=double_drop)

    def forward(self, x, mask=None, **kwargs):
        return self.ff(self.mhra(x, mask=mask, **kwargs))


# ---------------------------------------------------------------------------------------------
# in-repo part: masks, MusicTransformerXL, BeatPositionEncoder  (deep_music_genre.py:1577-1665)
# ---------------------------------------------------------------------------------------------
def window_mask(x_len, device, m_len=0, size=(1, 1)):
    "deep_music_genre.py:1577-1584"
    win_size, k = size
    mem_mask = torch.zeros((x_len, m_len), device=device)
    tri_mask = torch.triu(torch.ones((x_len // win_size + 1, x_len // win_size + 1), device=device), diagonal=k)
    wm = tri_mask.repeat_interleave(win_size, dim=0).repeat_interleave(win_size, dim=1)[:x_len, :x_len]
    if x_len:
        wm[..., 0] = 0
    mask = torch.cat((mem_mask, wm), dim=1)[None, None]
    return mask.bool()


def rand_window_mask(x_len, m_len, device, max_size=None, p=0.2, is_eval=False, rng=np.random):
    "deep_music_genre.py:1586-1590"
    if is_eval or rng.rand() >= p or max_size is None:
        win_size, k = (1, 1)
    else:
        win_size, k = (rng.randint(0, max_size) + 1, 0)
    return window_mask(x_len, device, m_len, size=(win_size, k))


class BeatPositionEncoder(nn.Module):
    "deep_music_genre.py:1651-1665"
    def __init__(self, emb_sz, beat_len=32, max_bar_len=1024):
        super().__init__()
        self.beat_len, self.max_bar_len = beat_len, max_bar_len
        self.beat_enc = nn.Embedding(beat_len, emb_sz, padding_idx=0)
        self.bar_enc = nn.Embedding(max_bar_len, emb_sz, padding_idx=0)

    def forward(self, pos):
        beat_enc = self.beat_enc(pos % self.beat_len)
        bar_pos = pos // self.beat_len % self.max_bar_len
        bar_pos[bar_pos >= self.max_bar_len] = self.max_bar_len - 1
        return beat_enc + self.bar_enc(bar_pos)


class MusicTransformerXL(nn.Module):
    """fastai TransformerXL.__init__/reset/select_hidden/_update_mems + the reference's forward override
    (deep_music_genre.py:1603-1647)."""
    def __init__(self, vocab_sz, ctx_len, n_layers, n_heads, d_model, d_head, d_inner, resid_p=0., attn_p=0.,
                 ff_p=0., embed_p=0., bias=False, scale=True, act='relu', double_drop=True, mask=True, mem_len=0,
                 encode_position=True, mask_steps=1, **unused):
        super().__init__()
        self.encoder = nn.Embedding(vocab_sz, d_model)
        self.pos_enc = PositionalEncoding(d_model)
        self.drop_emb = nn.Dropout(embed_p)
        self.u = nn.Parameter(torch.Tensor(n_heads, 1, d_head))
        self.v = nn.Parameter(torch.Tensor(n_heads, 1, d_head))
        self.mem_len, self.n_layers, self.d_model, self.mask = mem_len, n_layers, d_model, mask
        self.init = False
        self.layers = nn.ModuleList([DecoderLayer(n_heads, d_model, d_head, d_inner, resid_p=resid_p, attn_p=attn_p,
                                                  ff_p=ff_p, bias=bias, scale=scale, act=act, double_drop=double_drop)
                                     for _ in range(n_layers)])
        self.encode_position = encode_position
        if self.encode_position:
            self.beat_enc = BeatPositionEncoder(d_model)
        self.mask_steps = mask_steps

    def reset(self):
        "fastai: hidden = [empty 1-D tensor] * (n_layers + 1)"
        self.hidden = [next(self.parameters()).data.new(0) for _ in range(self.n_layers + 1)]

    def _update_mems(self, hids):
        if not getattr(self, 'hidden', False):
            return
        assert len(hids) == len(self.hidden), 'len(hids) != len(self.hidden)'
        with torch.no_grad():
            for i in range(len(hids)):
                cat = hids[i] if self.hidden[i].dim() == 1 else torch.cat([self.hidden[i], hids[i]], dim=1)
                self.hidden[i] = cat[:, -self.mem_len:].detach()

    def select_hidden(self, idxs):
        self.hidden = [h[idxs] for h in self.hidden]

    def forward(self, x):
        if self.mem_len > 0 and not self.init:
            self.reset()
            self.init = True
        benc = 0
        if self.encode_position:
            x, pos = x['x'], x['pos']
            benc = self.beat_enc(pos)
        bs, x_len = x.size()
        inp = self.drop_emb(self.encoder(x) + benc)
        m_len = self.hidden[0].size(1) if hasattr(self, 'hidden') and len(self.hidden[0].size()) > 1 else 0
        seq_len = m_len + x_len
        mask = rand_window_mask(x_len, m_len, inp.device, max_size=self.mask_steps,
                                is_eval=not self.training) if self.mask else None
        if m_len == 0 and mask is not None:
            mask[..., 0, 0] = 0
        hids = []
        pos = torch.arange(seq_len - 1, -1, -1, device=inp.device, dtype=inp.dtype)
        pos_enc = self.pos_enc(pos)
        hids.append(inp)
        for i, layer in enumerate(self.layers):
            mem = self.hidden[i] if self.mem_len > 0 else None
            inp = layer(inp, r=pos_enc, u=self.u, v=self.v, mask=mask, mem=mem)
            hids.append(inp)
        core_out = inp[:, -x_len:]
        if self.mem_len > 0:
            self._update_mems(hids)
        return (self.hidden if self.mem_len > 0 else [core_out]), [core_out]


class RNNDropout(nn.Module):
    "fastai RNNDropout: one mask per (batch, feature), shared over time; identity in eval."
    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def forward(self, x):
        if not self.training or self.p == 0.:
            return x
        m = x.new_empty(x.size(0), 1, x.size(2)).bernoulli_(1 - self.p).div_(1 - self.p)
        return x * m


class LinearDecoder(nn.Module):
    "fastai awd_lstm.LinearDecoder (tied head)."
    initrange = 0.1

    def __init__(self, n_out, n_hid, output_p, tie_encoder=None, bias=True):
        super().__init__()
        self.decoder = nn.Linear(n_hid, n_out, bias=bias)
        self.decoder.weight.data.uniform_(-self.initrange, self.initrange)
        self.output_dp = RNNDropout(output_p)
        if bias:
            self.decoder.bias.data.zero_()
        if tie_encoder is not None:
            self.decoder.weight = tie_encoder.weight

    def forward(self, input):
        raw_outputs, outputs = input
        output = self.output_dp(outputs[-1])
        decoded = self.decoder(output)
        return decoded, raw_outputs, outputs


class SequentialRNN(nn.Sequential):
    "fastai SequentialRNN: passes reset() to children."
    def reset(self):
        for c in self.children():
            if hasattr(c, 'reset'):
                c.reset()


def init_transformer(m):
    "fastai init_transformer (applied with model.apply)."
    classname = m.__class__.__name__
    if classname.find('Linear') != -1:
        if hasattr(m, 'weight') and m.weight is not None:
            nn.init.normal_(m.weight, 0., 0.02)
        if hasattr(m, 'bias') and m.bias is not None:
            nn.init.constant_(m.bias, 0.)
    elif classname.find('LayerNorm') != -1:
        if hasattr(m, 'weight') and m.weight is not None:
            nn.init.normal_(m.weight, 1., 0.02)
        if hasattr(m, 'bias') and m.bias is not None:
            nn.init.constant_(m.bias, 0.)
    elif classname.find('TransformerXL') != -1:
        if hasattr(m, 'u'):
            nn.init.normal_(m.u, 0., 0.02)
        if hasattr(m, 'v'):
            nn.init.normal_(m.v, 0., 0.02)


def get_language_model(vocab_sz, config, drop_mult=1.):
    "fastai.text.learner.get_language_model for arch=MusicTransformerXL (deep_music_genre.py:1793)."
    config = dict(config)
    for k in config.keys():
        if k.endswith('_p'):
            config[k] *= drop_mult
    tie_weights, output_p, out_bias = map(config.pop, ['tie_weights', 'output_p', 'out_bias'])
    encoder = MusicTransformerXL(vocab_sz, **config)
    enc = encoder.encoder if tie_weights else None
    decoder = LinearDecoder(vocab_sz, config['d_model'], output_p, tie_encoder=enc, bias=out_bias)
    model = SequentialRNN(encoder, decoder)
    return model.apply(init_transformer)


def count_parameters(model):
    "The notebook's calc_net_weight_count: trainable parameters, shared tensors counted once."
    seen, n = set(), 0
    for p in model.parameters():
        if p.requires_grad and id(p) not in seen:
            seen.add(id(p))
            n += p.numel()
    return n
