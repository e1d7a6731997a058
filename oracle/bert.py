"""Oracle: masked-BERT remix encoder + head (the `msk` branch of MultiTransformer).  TEST INFRASTRUCTURE.

Follows ``deep_music_remix.py:1851-2104``.  Only what ``MultiTransformer.forward`` runs for
``{'msk': {'x','pos'}}`` is restated: ``head(encoder(x, pos))`` (``:1880-1881``).  For that task every
``MTEncoderBlock`` returns right after ``mha1`` (``:2015-2016``): no ``mha2``, no FFN - but their
parameters exist in the reference module tree, so they are created here too (state-dict key parity,
``strict=False`` loading) and left unused.

Parity status: logits unpinned (see ``oracle/__init__.py``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .txl import PositionalEncoding, RNNDropout, _line_shift, feed_forward, init_transformer, music_config


def multitask_config():
    "app_utils.py:55-63"
    c = music_config()
    c.update(encode_position=True, bias=True, enc_layers=10, dec_layers=10)
    del c['n_layers']
    return c


class TransformerEmbedding(nn.Module):
    "deep_music_remix.py:1910-1938"
    def __init__(self, vocab_size, emb_sz, embed_p=0., mem_len=512, beat_len=32, max_bar_len=1024, pad_idx=None):
        super().__init__()
        self.emb_sz, self.pad_idx = emb_sz, pad_idx
        self.embed = nn.Embedding(vocab_size, emb_sz, padding_idx=pad_idx)
        self.pos_enc = PositionalEncoding(emb_sz)
        self.beat_len, self.max_bar_len = beat_len, max_bar_len
        self.beat_enc = nn.Embedding(beat_len, emb_sz, padding_idx=0)
        self.bar_enc = nn.Embedding(max_bar_len, emb_sz, padding_idx=0)
        self.drop = nn.Dropout(embed_p)
        self.mem_len = mem_len

    def forward(self, inp, pos):
        beat_enc = self.beat_enc(pos % self.beat_len)
        bar_pos = pos // self.beat_len % self.max_bar_len
        bar_pos[bar_pos >= self.max_bar_len] = self.max_bar_len - 1
        bar_enc = self.bar_enc(bar_pos)
        return self.drop(self.embed(inp) + beat_enc + bar_enc)

    def relative_pos_enc(self, emb):
        seq_len = emb.shape[1] + self.mem_len
        pos = torch.arange(seq_len - 1, -1, -1, device=emb.device, dtype=emb.dtype)
        return self.pos_enc(pos)


class MemMultiHeadRelativeAttentionKV(nn.Module):
    "deep_music_remix.py:2025-2104 (mem_len = 0 in the encoder -> mem_k/mem_v are the identity)"
    def __init__(self, n_heads, d_model, d_head=None, resid_p=0., attn_p=0., bias=True, scale=True, mem_len=512,
                 r_mask=True):
        super().__init__()
        d_head = d_head if d_head is not None else d_model // n_heads
        self.n_heads, self.d_head, self.scale = n_heads, d_head, scale
        assert d_model == d_head * n_heads
        self.q_wgt = nn.Linear(d_model, n_heads * d_head, bias=bias)
        self.k_wgt = nn.Linear(d_model, n_heads * d_head, bias=bias)
        self.v_wgt = nn.Linear(d_model, n_heads * d_head, bias=bias)
        self.drop_att, self.drop_res = nn.Dropout(attn_p), nn.Dropout(resid_p)
        self.ln = nn.LayerNorm(d_model)
        self.r_attn = nn.Linear(d_model, n_heads * d_head, bias=bias)
        self.r_mask = r_mask
        self.mem_len = mem_len

    def forward(self, q, k=None, v=None, r=None, g_u=None, g_v=None, mask=None):
        if k is None: k = q
        if v is None: v = q
        return self.ln(q + self.drop_res(self._apply_attention(q, k, v, r, g_u, g_v, mask=mask)))

    def _apply_attention(self, q, k, v, r=None, g_u=None, g_v=None, mask=None):
        assert self.mem_len == 0, 'oracle restates the encoder (mem_len=0) only'
        bs, x_len, seq_len = q.size(0), q.size(1), k.size(1)
        wq, wk, wv = self.q_wgt(q), self.k_wgt(k), self.v_wgt(v)
        wq = wq[:, -x_len:]
        wq, wk, wv = map(lambda t: t.view(bs, t.size(1), self.n_heads, self.d_head), (wq, wk, wv))
        wq, wk, wv = wq.permute(0, 2, 1, 3), wk.permute(0, 2, 3, 1), wv.permute(0, 2, 1, 3)
        wkr = self.r_attn(r[-seq_len:])
        wkr = wkr.view(seq_len, self.n_heads, self.d_head)
        wkr = wkr.permute(1, 2, 0)
        AC = torch.matmul(wq + g_u, wk)
        BD = _line_shift(torch.matmul(wq + g_v, wkr), mask=self.r_mask)
        attn_score = AC + BD
        if self.scale:
            attn_score = attn_score.mul_(1 / (self.d_head ** 0.5))
        if mask is not None:
            mask = mask[..., -seq_len:].bool()
            attn_score = attn_score.float().masked_fill(mask, -float('inf')).type_as(attn_score)
        attn_prob = self.drop_att(F.softmax(attn_score, dim=-1))
        attn_vec = torch.matmul(attn_prob, wv)
        return attn_vec.permute(0, 2, 1, 3).contiguous().view(bs, x_len, -1)


class MTEncoderBlock(nn.Module):
    "deep_music_remix.py:2000-2017"
    def __init__(self, n_heads, d_model, d_head, d_inner, resid_p=0., attn_p=0., ff_p=0., bias=True, scale=True,
                 double_drop=True, mem_len=512, mha2_mem_len=0, **kwargs):
        super().__init__()
        cls = MemMultiHeadRelativeAttentionKV
        self.mha1 = cls(n_heads, d_model, d_head, resid_p=resid_p, attn_p=attn_p, bias=bias, scale=scale,
                        mem_len=mem_len, r_mask=False)
        self.mha2 = cls(n_heads, d_model, d_head, resid_p=resid_p, attn_p=attn_p, bias=bias, scale=scale,
                        mem_len=mha2_mem_len, r_mask=True)
        self.ff = feed_forward(d_model, d_inner, ff_p=ff_p, double_drop=double_drop)   # act not passed (:2009) -> ReLU

    def forward(self, enc_lm, enc_msk, r=None, g_u=None, g_v=None, msk_mask=None, lm_mask=None):
        y_lm = self.mha1(enc_lm, enc_lm, enc_lm, r, g_u, g_v, mask=lm_mask)
        if enc_msk is None:
            return y_lm
        raise NotImplementedError('decoder / s2s branch is out of scope (SURVEY.md 2.1)')


class MTEncoder(nn.Module):
    "deep_music_remix.py:1959-1998 (is_decoder=False)"
    def __init__(self, embed, n_hid, n_layers, n_heads, d_model, d_head, d_inner, resid_p=0., attn_p=0., ff_p=0.,
                 bias=True, scale=True, act='relu', double_drop=True, mem_len=512, is_decoder=False, mask_steps=1,
                 mask_p=0.3, **kwargs):
        super().__init__()
        assert not is_decoder
        self.embed = embed
        self.u = nn.Parameter(torch.Tensor(n_heads, 1, d_head))
        self.v = nn.Parameter(torch.Tensor(n_heads, 1, d_head))
        self.n_layers, self.d_model = n_layers, d_model
        self.layers = nn.ModuleList([MTEncoderBlock(n_heads, d_model, d_head, d_inner, resid_p=resid_p, attn_p=attn_p,
                                                    ff_p=ff_p, bias=bias, scale=scale, act=act,
                                                    double_drop=double_drop, mem_len=mem_len)
                                     for _ in range(n_layers)])
        nn.init.normal_(self.u, 0., 0.02)
        nn.init.normal_(self.v, 0., 0.02)

    def forward(self, x_lm, lm_pos, msk_emb=None):
        lm_emb = self.embed(x_lm, lm_pos)
        pos_enc = self.embed.relative_pos_enc(lm_emb)
        for layer in self.layers:
            lm_emb = layer(lm_emb, msk_emb, lm_mask=None, r=pos_enc, g_u=self.u, g_v=self.v)
        return lm_emb


class MTLinearDecoder(nn.Module):
    "deep_music_remix.py:1940-1955"
    initrange = 0.1

    def __init__(self, n_hid, n_out, output_p, tie_encoder=None, out_bias=True, **kwargs):
        super().__init__()
        self.decoder = nn.Linear(n_hid, n_out, bias=out_bias)
        self.decoder.weight.data.uniform_(-self.initrange, self.initrange)
        self.output_dp = RNNDropout(output_p)
        if out_bias:
            self.decoder.bias.data.zero_()
        if tie_encoder is not None:
            self.decoder.weight = tie_encoder.weight

    def forward(self, input):
        return self.decoder(self.output_dp(input))


class MultiTransformer(nn.Module):
    "deep_music_remix.py:1864-1902, mask task only (the decoder module is out of scope and not built)."
    def __init__(self, encoder, head, mem_len):
        super().__init__()
        self.encoder, self.head = encoder, head
        self.default_mem_len = mem_len

    def forward(self, inp):
        outputs = {}
        msk = inp.get('msk')
        if msk is not None:
            outputs['msk'] = self.head(self.encoder(msk['x'], msk['pos']))
        for key in ('lm', 's2f', 'f2s'):
            if inp.get(key) is not None:
                raise NotImplementedError(f'task {key!r} is out of scope (SURVEY.md 2.1)')
        return outputs

    def reset(self):
        pass


def get_multitask_model(vocab_size, config, drop_mult=1., pad_idx=None):
    "deep_music_remix.py:1851-1862 (encoder + head)."
    config = dict(config)
    for k in config.keys():
        if k.endswith('_p'):
            config[k] *= drop_mult
    n_hid = config['d_model']
    mem_len = config.pop('mem_len')
    embed = TransformerEmbedding(vocab_size, n_hid, embed_p=config['embed_p'], mem_len=mem_len, pad_idx=pad_idx)
    encoder = MTEncoder(embed, n_hid, n_layers=config['enc_layers'], mem_len=0, **config)
    head = MTLinearDecoder(n_hid, vocab_size, tie_encoder=embed.embed, **config)
    model = MultiTransformer(encoder, head, mem_len=mem_len)
    return model.apply(init_transformer)
