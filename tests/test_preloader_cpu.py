"""SURVEY.md section 8 f4 (training data feed): the oracle restatement of MusicPreloader / mask_tfm against batches produced by
the reference's OWN source (tests/golden/make_preloader_golden.py executes deep_music_genre.py:1001-1125 and
deep_music_remix.py:1208-1223 with the fastai base class stubbed), and the host logic of the product mirror."""
import os

import numpy as np
import pytest
import torch

from oracle import preloader as opl

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preloader_golden.npz'))
NOTE_RANGE = (12, 140)
CASES = {'a': dict(bs=4, bptt=16, shuffle=True, transpose_range=(0, 12), encode_position=True, backwards=False),
         'b': dict(bs=3, bptt=20, shuffle=False, transpose_range=None, encode_position=False, backwards=False),
         'c': dict(bs=5, bptt=8, shuffle=True, transpose_range=(0, 24), encode_position=False, backwards=True)}


def golden_items(name):
    lens, data, pos = GOLD[f'{name}_lens'], GOLD[f'{name}_data'], GOLD[f'{name}_pos']
    off = np.concatenate([[0], np.cumsum(lens)])
    return [opl.Item(data[off[i]:off[i + 1]], pos[off[i]:off[i + 1]]) for i in range(len(lens))]


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_preloader_equals_reference_source(name):
    "same seeds, same RNG call order as the reference -> identical permutations, transposes and batches over two epochs"
    c = CASES[name]
    torch.manual_seed(10); np.random.seed(10)
    pl = opl.MusicPreloader(golden_items(name), NOTE_RANGE, **c)
    for epoch in range(2):
        pl.on_epoch_begin()
        xs, ys = [], []
        for k in range(len(pl)):
            x, y = pl[k]
            xs.append(np.array(x)); ys.append(np.array(y))
        assert np.array_equal(pl.idx.idx, GOLD[f'{name}_perm{epoch}'])
        if c['transpose_range'] is not None:
            assert np.array_equal(pl.transpose_values.numpy(), GOLD[f'{name}_tv{epoch}'])
        assert np.array_equal(np.stack(xs), GOLD[f'{name}_x{epoch}'])
        assert np.array_equal(np.stack(ys), GOLD[f'{name}_y{epoch}'])


def test_oracle_preloader_streams_are_contiguous():
    "row j of batch b+1 starts with the token row j of batch b ended with (overlap 1), and y is x shifted by one"
    rng = np.random.default_rng(0)
    items = [opl.Item(rng.integers(0, 324, n), np.arange(n)) for n in rng.integers(3, 50, 13)]
    pl = opl.MusicPreloader(items, NOTE_RANGE, bs=4, bptt=12, encode_position=False)
    prev = None
    for x, y in pl.batches():
        assert np.array_equal(x[:, 1:], y[:, :-1])
        if prev is not None:
            assert np.array_equal(x[:, 0], prev[:, -1])
        prev = y


def test_oracle_mask_tfm_equals_reference_source():
    x = torch.from_numpy(GOLD['mask_x'])
    mx, my = opl.mask_tfm(x, x.clone(), (12, 301), 4, 1, p=0.3, rand=torch.from_numpy(GOLD['mask_rand']),
                          wrong=torch.from_numpy(GOLD['mask_wrong']))
    assert np.array_equal(mx.numpy(), GOLD['mask_out_x']) and np.array_equal(my.numpy(), GOLD['mask_out_y'])
    torch.manual_seed(77)                                   # and with the reference's own draws
    mx, my = opl.mask_tfm(x, x.clone(), (12, 301), 4, 1, p=0.3)
    assert np.array_equal(mx.numpy(), GOLD['mask_out_x']) and np.array_equal(my.numpy(), GOLD['mask_out_y'])


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
def test_product_mirror_host_bookkeeping_equals_oracle(name):
    "the mirror's per-epoch host state (permutation, transposes, row cursors) is the reference's, draw for draw"
    from deepmusicgeneration_b200.preloader import MusicPreloader
    c = dict(CASES[name])
    if c['backwards'] and c['encode_position']:
        pytest.skip('fails in the reference')
    items = golden_items(name)
    torch.manual_seed(10); np.random.seed(10)
    ref = opl.MusicPreloader(items, NOTE_RANGE, **c)
    torch.manual_seed(10); np.random.seed(10)
    got = MusicPreloader(items, note_range=NOTE_RANGE, **c)
    for epoch in range(2):
        torch.manual_seed(20 + epoch); np.random.seed(20 + epoch)       # both draw from the global generators: same state for each
        ref.on_epoch_begin()
        torch.manual_seed(20 + epoch); np.random.seed(20 + epoch)
        got.on_epoch_begin()
        assert len(ref) == len(got)
        assert np.array_equal(ref.idx.idx, got.idx.idx) and np.array_equal(ref.ro, got.ro) and np.array_equal(ref.ri, got.ri)
        if c['transpose_range'] is not None:
            assert torch.equal(ref.transpose_values, got.transpose_values)
        for k in range(len(ref)):                  # advance the oracle so that the next epoch's RNG position matches
            ref[k]


def test_product_mirror_has_no_cpu_path():
    from deepmusicgeneration_b200.preloader import MusicPreloader
    if torch.cuda.is_available():
        pytest.skip('needs a machine without a GPU')
    pl = MusicPreloader(golden_items('b'), note_range=NOTE_RANGE, bs=3, bptt=20, encode_position=False)
    with pytest.raises(RuntimeError, match='no CPU path'):
        pl.next_batch()
