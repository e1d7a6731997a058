"""SURVEY.md section 8 f4 on the GPU: dmg_preload_fill / dmg_mask_tfm through the C ABI against the oracle (bit-exact, int64)."""
import os

import numpy as np
import pytest
import torch

from oracle import preloader as opl
from deepmusicgeneration_b200.preloader import MusicPreloader, mask_tfm

pytestmark = pytest.mark.gpu
NOTE_RANGE = (12, 140)
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preloader_golden.npz'))


def random_items(seed, n, lo, hi):
    rng = np.random.default_rng(seed)
    return [opl.Item(rng.integers(0, 324, L), np.cumsum(rng.integers(0, 5, L))) for L in rng.integers(lo, hi, n)]


def run_epochs(items, epochs=2, **c):
    torch.manual_seed(3); np.random.seed(3)
    ref = opl.MusicPreloader(items, NOTE_RANGE, **c)
    torch.manual_seed(3); np.random.seed(3)
    got = MusicPreloader(items, note_range=NOTE_RANGE, **c)
    for epoch in range(epochs):
        torch.manual_seed(20 + epoch); np.random.seed(20 + epoch)       # both draw from the global generators: same state for each
        rb = list(ref.batches())
        torch.manual_seed(20 + epoch); np.random.seed(20 + epoch)
        gb = list(got)
        assert len(rb) == len(gb) == got.n_batches
        for (rx, ry), (gx, gy) in zip(rb, gb):
            if c.get('encode_position', True):
                assert np.array_equal(rx['x'], gx['x'].cpu().numpy()) and np.array_equal(rx['pos'], gx['pos'].cpu().numpy())
            else:
                assert np.array_equal(rx, gx.cpu().numpy())
            assert np.array_equal(ry, gy.cpu().numpy())


@pytest.mark.parametrize('c', [
    dict(bs=4, bptt=16, shuffle=True, transpose_range=(0, 12), encode_position=True),
    dict(bs=3, bptt=20, shuffle=False, transpose_range=None, encode_position=False),
    dict(bs=5, bptt=8, shuffle=True, transpose_range=(0, 24), encode_position=False, backwards=True),
    dict(bs=32, bptt=512, shuffle=True, transpose_range=(0, 12), encode_position=False),          # C3 geometry
    dict(bs=2, bptt=300, shuffle=True, transpose_range=(0, 12), encode_position=True),            # rows spanning > 64 items
])
def test_preloader_batches_equal_oracle(c):
    short = c['bptt'] == 300
    items = random_items(7, 400 if c['bptt'] >= 300 else 23, 1 if short else 3, 4 if short else 90)
    run_epochs(items, **c)


def test_preloader_reproduces_reference_source_batches():
    "the fixture written by the reference's own code (tests/golden/make_preloader_golden.py), case a"
    lens, data, pos = GOLD['a_lens'], GOLD['a_data'], GOLD['a_pos']
    off = np.concatenate([[0], np.cumsum(lens)])
    items = [opl.Item(data[off[i]:off[i + 1]], pos[off[i]:off[i + 1]]) for i in range(len(lens))]
    torch.manual_seed(10); np.random.seed(10)
    pl = MusicPreloader(items, note_range=NOTE_RANGE, bs=4, bptt=16, shuffle=True, transpose_range=(0, 12), encode_position=True)
    for epoch in range(2):
        xs, ys = [], []
        for x, y in pl:
            xs.append(torch.stack([x['x'], x['pos']], -1).cpu().numpy()); ys.append(y.cpu().numpy())
        gx, gy = GOLD[f'a_x{epoch}'], GOLD[f'a_y{epoch}']                 # [items, bptt, 2], item k = batch k // bs, row k % bs
        assert np.array_equal(np.concatenate(xs), gx) and np.array_equal(np.concatenate(ys), gy[..., 0])


def test_preloader_rank_rows_tile_the_global_batch():
    "world 2: the two ranks' rows are rows 0..bs-1 and bs..2bs-1 of the world-1 loader with the doubled batch (deep_music_genre.py:1021)"
    items = random_items(11, 60, 5, 70)
    torch.manual_seed(5); np.random.seed(5)
    whole = MusicPreloader(items, note_range=NOTE_RANGE, bs=8, bptt=32, shuffle=True, transpose_range=(0, 12), encode_position=False)
    parts = []
    for r in range(2):
        torch.manual_seed(5); np.random.seed(5)
        parts.append(MusicPreloader(items, note_range=NOTE_RANGE, bs=4, bptt=32, shuffle=True, transpose_range=(0, 12), encode_position=False,
                                    world=2, rank=r))
    def epoch(pl):
        torch.manual_seed(6); np.random.seed(6)
        return list(pl)
    for (x, y), (x0, y0), (x1, y1) in zip(epoch(whole), epoch(parts[0]), epoch(parts[1])):
        assert torch.equal(x, torch.cat([x0, x1])) and torch.equal(y, torch.cat([y0, y1]))


def test_mask_tfm_equals_oracle_on_its_own_draws():
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 324, (64, 1024), generator=g).cuda()
    y = x.clone()
    mx, my, rand, wrong = mask_tfm((x, y), (12, 301), 4, 1, p=0.3, seed=1234, return_draws=True)
    r2 = rand.clone(); r2[(x < 12) | (x >= 301)] = 1.0
    ww = (r2 > 0.3 * .8) & (r2 <= 0.3 * .9)
    ox, oy = opl.mask_tfm(x, y, (12, 301), 4, 1, p=0.3, rand=rand, wrong=wrong[ww])
    assert torch.equal(mx, ox) and torch.equal(my, oy)
    # the statistics the reference's comment promises: p of the in-range tokens selected, 80 / 10 / 10 split
    inr = (x >= 12) & (x < 301)
    sel = (my != 1) & inr
    frac = sel.sum().item() / inr.sum().item()
    assert abs(frac - 0.3) < 0.01
    assert abs(((mx == 4) & inr).sum().item() / sel.sum().item() - 0.8) < 0.02
    assert abs(rand.mean().item() - 0.5) < 0.01 and rand.min().item() >= 0.0 and rand.max().item() < 1.0
    assert torch.equal(mx[~inr], x[~inr])


def test_fit_one_cycle_consumes_the_device_preloader():
    "music_model_learner(...).fit_one_cycle(epochs, lr, MusicPreloader): the notebook's training call end to end on the GPU"
    from deepmusicgeneration_b200.codec import MusicDataBunch
    from deepmusicgeneration_b200.learner import music_model_learner
    from oracle import txl
    cfg = dict(txl.default_config(), n_layers=2, d_model=128, n_heads=2, d_head=64, d_inner=256, mem_len=64, encode_position=False)
    data = MusicDataBunch.empty('')
    learn = music_model_learner(data, config=cfg, encode_position=False, dtype='bf16', max_batch=4, max_seq=64, keep_hidden=False, seed=0)
    rng = np.random.default_rng(0)
    pattern = rng.integers(12, 300, 48)
    items = [opl.Item(np.tile(pattern, 6)[:int(n)], np.arange(int(n))) for n in rng.integers(120, 288, 12)]
    pl = MusicPreloader(items, data.vocab, bs=4, bptt=64, shuffle=True, transpose_range=None, encode_position=False)
    first = {}
    def cb(i, tr):
        if i == 0: first.update(tr.losses())
    last = learn.fit_one_cycle(6, 3e-3, pl, callback=cb)
    assert last['ce'] < 0.7 * first['ce'], (first, last)
