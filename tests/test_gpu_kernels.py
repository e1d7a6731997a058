"""GPU unit tests of single kernels through the C ABI: tcgen05 GEMM, SIMT GEMM, the fused sampler."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import codec as ocodec, sampling as osamp

pytestmark = pytest.mark.gpu
V = 324


def _lib():
    from deepmusicgeneration_b200 import _lib
    return _lib, _lib.load()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@pytest.mark.parametrize('M,N,K,gelu,out_bf16,bias', [
    (256, 1536, 512, 0, 0, False),     # decode QKV
    (256, 512, 2048, 0, 0, True),      # decode FF2
    (256, 2048, 512, 1, 1, True),      # decode FF1 (+GeLU, bf16 out)
    (256, 324, 512, 0, 0, True),       # tied head, N not a tile multiple
    (1, 512, 512, 0, 0, False),        # single stream
    (131, 324, 768, 0, 0, True),       # ragged M, K = 12 heads x 64
    (1000, 2048, 512, 1, 1, True),     # M > 512 -> 128x128 tiles
    (4096, 1536, 512, 0, 0, False),    # prefill-sized
    (640, 512, 3072, 0, 0, True),      # long K (48 k-blocks through a 6-stage ring)
    (256, 512, 3072, 0, 0, True),      # split-K 8 x 6 k-blocks
    (64, 2048, 768, 1, 1, True),       # split-K 4 x 3 k-blocks, GeLU + bf16 out
    (300, 324, 2048, 0, 0, True),      # 3 M-tiles, ragged N, split-K 8
])
def test_gemm_tcgen05_vs_fp32(M, N, K, gelu, out_bf16, bias):
    L, lib = _lib()
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N)
    a = (torch.randn(M, K, device='cuda', generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device='cuda', generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device='cuda', generator=g) if bias else None
    ref = a.float() @ w.float().t()
    if bias: ref = ref + b
    if gelu:
        ref = 0.5 * ref * (1 + torch.tanh(0.7978845608028654 * (ref + 0.044715 * ref ** 3)))
    for backend in (L.GEMM_AUTO, L.GEMM_TC_TILE, L.GEMM_SIMT):   # AUTO = cluster split-K when M <= 512
        c = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16 if out_bf16 else torch.float32)
        rc = lib.dmg_gemm_bf16(_p(a), _p(w), _p(b), _p(c), M, N, K, gelu, out_bf16, backend, C.c_void_p(0))
        assert rc == 0, lib.dmg_last_error()
        torch.cuda.synchronize()
        err = (c.float() - ref).abs().max().item()
        tol = 2e-2 * max(1.0, ref.abs().max().item()) if out_bf16 else 2e-3 * max(1.0, ref.abs().max().item())
        assert not torch.isnan(c.float()).any(), f'backend {backend}: NaN left in output (tile not written)'
        assert err < tol, f'backend {backend}: max err {err} (tol {tol})'


@pytest.fixture(scope='module')
def tiny_engine():
    from deepmusicgeneration_b200.model import get_multitask_model
    cfg = dict(d_model=128, n_heads=2, d_head=64, d_inner=256, enc_layers=1, mem_len=512, bias=True)
    return get_multitask_model(V, cfg, dtype='f32', max_batch=1, max_seq=8, seed=0)


def _run_sampler(engine, logits, prev, rc, temperatures, top_k, top_p, offset=0, seed=1):
    from deepmusicgeneration_b200 import _lib as L
    from deepmusicgeneration_b200.codec import MusicVocab
    from deepmusicgeneration_b200.learner import sampler_params, vocab_layout
    vocab = MusicVocab.create()
    vl = vocab_layout(vocab)
    params = sampler_params(vocab, 1, temperatures, 0, top_k, top_p, None, flags=L.SAMPLE_REMIX_FILTER, seed=seed)
    n = logits.shape[0]
    out = torch.zeros(n, dtype=torch.int32, device='cuda'); nc = torch.zeros(n, dtype=torch.int32, device='cuda')
    e = engine._e
    rcode = e.lib.dmg_sample_logits(e.h, _p(logits), _p(prev), _p(rc), n, C.byref(vl), C.byref(params), offset, _p(out), _p(nc),
                                    C.c_void_p(0))
    assert rcode == 0, e.lib.dmg_last_error()
    torch.cuda.synchronize()
    return out.cpu(), nc.cpu()


def _oracle_filtered(logits_row, prev, rc, temperatures, top_k, top_p):
    "predict_mask's sampling front half on the CPU (deep_music_remix.py:2586-2600) -> final probs"
    v = ocodec.MusicVocab.create()
    logits = logits_row.clone()
    temperature = temperatures[0] if v.is_duration_or_pad(prev) else temperatures[1]
    temperature += max(0, np.log((rc + 1) / 4) / 5) * temperature
    if temperature != 1.: logits = logits / temperature
    special = [v.bos_idx, v.sep_idx, v.stoi['xxni'], v.stoi['xxeos']] + [v.stoi[t] for t in ('xxelec', 'xxfolk', 'xxfunk', 'xxjazz', 'xxpop', 'xxrock')]
    logits[special] = -float('inf')
    logits = osamp.filter_invalid_indexes_remix(logits, prev, v, filter_value=-float('inf'))
    logits = osamp.top_k_top_p(logits, top_k=top_k, top_p=top_p)
    return torch.softmax(logits, -1)


def test_sampler_set_sizes_and_greedy_match_oracle(tiny_engine):
    g = torch.Generator().manual_seed(0)
    n = 64
    logits = torch.randn(n, V, generator=g) * 2
    v = ocodec.MusicVocab.create()
    prevs = [v.pad_idx, v.stoi['d4'], v.stoi['i0'], v.stoi['n60'], v.sep_idx, v.stoi['xxni'], v.stoi['d160'], v.stoi['n0']]
    prev = torch.tensor([prevs[i % len(prevs)] for i in range(n)], dtype=torch.int32)
    rc = torch.tensor([(i * 3) % 11 for i in range(n)], dtype=torch.int32)
    for top_k, top_p in ((1, 0.0), (20, 0.8), (40, 0.6), (0, 0.9), (5, 0.0), (0, 0.0), (400, 0.3)):
        out, nc = _run_sampler(tiny_engine, logits.cuda(), prev.cuda(), rc.cuda(), (1.2, 0.8), top_k, top_p)
        for i in range(n):
            probs = _oracle_filtered(logits[i], int(prev[i]), int(rc[i]), (1.2, 0.8), top_k, top_p)
            assert int(nc[i]) == int((probs > 0).sum()), (top_k, top_p, i)
            assert probs[int(out[i])] > 0, (top_k, top_p, i)
            if top_k == 1:
                assert int(out[i]) == int(probs.argmax())


def test_sampler_distribution_matches_oracle_probs(tiny_engine):
    g = torch.Generator().manual_seed(1)
    v = ocodec.MusicVocab.create()
    row = torch.randn(V, generator=g) * 1.5
    n = 20000
    logits = row[None].repeat(n, 1).contiguous()
    prev = torch.full((n,), v.stoi['i0'], dtype=torch.int32); rc = torch.zeros(n, dtype=torch.int32)
    out, _ = _run_sampler(tiny_engine, logits.cuda(), prev.cuda(), rc.cuda(), (1.0, 1.0), 12, 0.9, seed=123)
    probs = _oracle_filtered(row, v.stoi['i0'], 0, (1.0, 1.0), 12, 0.9)
    freq = torch.bincount(out.long(), minlength=V).float() / n
    assert (freq[probs == 0] == 0).all()
    assert (freq - probs).abs().max() < 4 * (probs.max() * (1 - probs.max()) / n) ** 0.5 + 2e-3
