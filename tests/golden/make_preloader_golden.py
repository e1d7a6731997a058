"""Generates tests/golden/preloader_golden.npz by EXECUTING the reference's own MusicPreloader / tfm_transpose / mask_tfm source
(deep_music_genre.py:1001-1125, :1541-1544; deep_music_remix.py:1208-1223) with the fastai base class stubbed out.

Run in the build container only (needs /root/reference): ``python tests/golden/make_preloader_golden.py``.
The reference module cannot be imported as a whole (fastai, music21 are not installed; np.int is gone), but these three
definitions are self-contained once ``Callback``, the typing names and ``num_distrib`` exist."""
import math, os, re, sys
import numpy as np, torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
np.int = int                                      # removed alias the reference still uses (:1052)


def grab(path, start_pat, end_pat):
    src = open(path).read()
    a = re.search(start_pat, src, re.M).start()
    b = re.search(end_pat, src[a:], re.M).start() + a
    return src[a:b]


ns = {'np': np, 'torch': torch, 'math': math, 'Callback': object, 'LabelList': object, 'Collection': list, 'Any': object,
      'num_distrib': lambda: 0}
exec(grab(os.path.join(REF, 'deep_music_genre.py'), r'^class MusicPreloader\(Callback\):', r'^def batch_position_tfm'), ns)
exec(grab(os.path.join(REF, 'deep_music_genre.py'), r'^def tfm_transpose', r'^def trim_to_beat'), ns)
exec(grab(os.path.join(REF, 'deep_music_remix.py'), r'^def mask_tfm\(', r'^def mask_lm_tfm_default'), ns)
MusicPreloader, tfm_transpose, mask_tfm = ns['MusicPreloader'], ns['tfm_transpose'], ns['mask_tfm']

NOTE_RANGE = (12, 140)                            # MusicVocab.note_range of the 324-entry vocabulary (tests/test_oracle_pins.py)


class Vocab:
    note_range = NOTE_RANGE


class Item:                                       # the slice of MusicItem (:1160-1249) the preloader touches
    def __init__(self, data, position): self.data, self.position, self.vocab = data, position, Vocab
    def __len__(self): return len(self.data)
    def transpose(self, interval): return Item(tfm_transpose(self.data, interval, self.vocab), self.position)


class Dataset:
    def __init__(self, items): self.x, self.vocab, self.item = items, Vocab, None
    def __len__(self): return len(self.x)


def make_items(rng, n, lo, hi):
    items = []
    for _ in range(n):
        L = int(rng.integers(lo, hi))
        data = rng.integers(0, 324, L).astype(np.int64)
        pos = np.cumsum(rng.integers(0, 5, L)).astype(np.int64)
        items.append(Item(data, pos))
    return items


out = {}
cases = {'a': dict(n=7, lo=5, hi=60, bs=4, bptt=16, shuffle=True, transpose_range=(0, 12), encode_position=True, backwards=False),
         'b': dict(n=11, lo=1, hi=9, bs=3, bptt=20, shuffle=False, transpose_range=None, encode_position=False, backwards=False),
         'c': dict(n=9, lo=3, hi=40, bs=5, bptt=8, shuffle=True, transpose_range=(0, 24), encode_position=False, backwards=True)}
for name, c in cases.items():
    rng = np.random.default_rng(hash(name) % 1000 + 1 if False else {'a': 1, 'b': 2, 'c': 3}[name])
    items = make_items(rng, c['n'], c['lo'], c['hi'])
    torch.manual_seed(10); np.random.seed(10)
    pl = MusicPreloader(Dataset(items), bs=c['bs'], bptt=c['bptt'], shuffle=c['shuffle'], transpose_range=c['transpose_range'],
                        encode_position=c['encode_position'], backwards=c['backwards'])
    for epoch in range(2):
        pl.on_epoch_begin()
        xs, ys = [], []
        for k in range(len(pl)):
            x, y = pl[k]
            xs.append(np.array(x)); ys.append(np.array(y))
        out[f'{name}_x{epoch}'], out[f'{name}_y{epoch}'] = np.stack(xs), np.stack(ys)
        out[f'{name}_perm{epoch}'] = pl.idx.idx.copy()
        if pl.transpose_values is not None: out[f'{name}_tv{epoch}'] = pl.transpose_values.numpy().copy()
    out[f'{name}_lens'] = np.array([len(i) for i in items])
    out[f'{name}_data'] = np.concatenate([i.data for i in items])
    out[f'{name}_pos'] = np.concatenate([i.position for i in items])

# mask_tfm: the reference draws rand / randint itself; record them by re-drawing under the same seed
g = np.random.default_rng(5)
x = torch.from_numpy(g.integers(0, 324, (6, 40)).astype(np.int64)); y = x.clone()
torch.manual_seed(77)
mx, my = mask_tfm((x, y), mask_range=(12, 301), mask_idx=4, pad_idx=1, p=0.3)
torch.manual_seed(77)
rand = torch.rand(x.shape)
r2 = rand.clone(); r2[x < 12] = 1.0; r2[x >= 301] = 1.0
n_wrong = int(((r2 > 0.3 * .8) & (r2 <= 0.3 * .9)).sum())
wrong = torch.randint(12, 301, [n_wrong])
out.update(mask_x=x.numpy(), mask_rand=rand.numpy(), mask_wrong=wrong.numpy(), mask_out_x=mx.numpy(), mask_out_y=my.numpy())
np.savez_compressed(os.path.join(HERE, 'preloader_golden.npz'), **out)
print('wrote', len(out), 'arrays')
