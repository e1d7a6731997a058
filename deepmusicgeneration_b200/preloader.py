"""Training data feed on the GPU: mirrors of the reference's ``MusicPreloader``
(deep_music_genre.py:1001-1125) and ``mask_tfm`` (deep_music_remix.py:1208-1223).

The tokenised corpus is uploaded once as a flat ragged array; every batch is ONE kernel launch
(``dmg_preload_fill``: all ``bs`` rows walk their items, apply the per-item random transpose and write x / y / pos).
Only the per-epoch bookkeeping stays on the host (vectorised numpy / torch); it draws from the same generators in the same
order as the reference, so the same seeds give the same shuffles and transposes: ``np.random.shuffle`` of the item permutation,
``torch.randint`` + ``torch.rand`` for the transpose values; the initial row cursors (ro, ri) are a searchsorted over the
cumulative item lengths.

    pl = MusicPreloader(items, vocab, bs=32, bptt=512, shuffle=True, transpose_range=(0, 12), encode_position=False)
    for x, y in pl:            # one epoch; x, y int64 [bs, bptt] on the GPU ({'x':…, 'pos':…}, y with encode_position)
        trainer.step(x, y)

``items``: anything with ``.data`` and ``.position`` (codec.MusicItem), or plain 1-D id arrays.  Under data parallelism the
reference multiplies ``bs`` by the world size (:1021); here ``world`` / ``rank`` select this rank's rows of that global batch.
There is no CPU path: the batches are produced by the CUDA library or not at all.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _ItemOrder:
    """The epoch's item order (the reference's ``CircularIndex``, deep_music_genre.py:1005-1014, as data instead of an indexer):
    ``idx`` is the permutation ``np.random.shuffle`` acts on - same RNG call, same stream as the reference - and ``visit_order``
    the order in which a row cursor walks the items (reversed when reading backwards, wrapping around)."""

    def __init__(self, length, forward):
        self.idx, self.forward = np.arange(length), forward

    def __len__(self):
        return len(self.idx)

    def shuffle(self):
        np.random.shuffle(self.idx)

    def visit_order(self):
        return self.idx if self.forward else self.idx[::-1]


class MusicPreloader:
    "Mirror of deep_music_genre.py:1001-1125; iterate it to get the batches of one epoch."

    def __init__(self, items, vocab=None, bs=32, bptt=70, backwards=False, shuffle=False, y_offset=1, transpose_range=None,
                 transpose_p=0.5, encode_position=True, note_range=None, world=1, rank=0, device=None):
        if note_range is None:
            note_range = vocab.note_range if vocab is not None else (0, 0)
        self.note_range = (int(note_range[0]), int(note_range[1]))
        datas = [np.asarray(getattr(it, 'data', it), dtype=np.int64) for it in items]
        self.lengths = np.array([len(d) for d in datas])
        assert len(datas) > 0 and (self.lengths > 0).all(), 'MusicPreloader: empty dataset / empty item'
        self.n_items = len(datas)
        self.local_bs, self.world, self.rank = bs, max(1, world), rank
        self.bs = bs * self.world                                                    # :1021
        self.bptt, self.shuffle, self.backwards, self.y_offset = bptt, shuffle, backwards, y_offset
        self.transpose_range, self.transpose_p, self.encode_position = transpose_range, transpose_p, encode_position
        if backwards and encode_position:
            raise ValueError('MusicPreloader: backwards=True with encode_position=True fails in the reference as well '
                             '(fill_row sizes the copy with row.size, deep_music_genre.py:1120)')
        self.device = None if device is None else torch.device(device)       # resolved at the first batch (current CUDA device)
        self._host_tokens = np.concatenate(datas)
        self._host_positions = None
        if encode_position:
            self._host_positions = np.concatenate([np.asarray(it.position, dtype=np.int64) for it in items])
            assert len(self._host_positions) == len(self._host_tokens), 'MusicPreloader: position / data length mismatch'
        self._dev = None
        self.totalToks, self.ite_len, self.idx = 0, None, None
        self.bptt_len = self.bptt
        self.allocate_buffers()

    # ---- host bookkeeping ------------------------------------------------------------------------------------------
    # Same results as the reference's loops (tests/test_preloader_cpu.py checks them against batches the reference's own
    # source produced); the random draws keep the reference's call order so equal seeds give equal epochs.
    def __len__(self):                                                               # items per epoch = bs * batches (:1032-1037)
        if self.ite_len is None:
            self.totalToks = self.lengths.sum()
            self.ite_len = self.bs * int(math.ceil(self.totalToks / (self.bptt * self.bs)))
        return self.ite_len

    @property
    def n_batches(self):
        return len(self) // self.bs

    def allocate_buffers(self):
        if self.ite_len is None:
            len(self)
        self.idx = _ItemOrder(self.n_items, not self.backwards)
        self.ro = np.zeros(self.bs, dtype=np.int64)
        self.ri = np.zeros(self.bs, dtype=np.int64)
        self.transpose_values = self.get_random_transpose_values()

    def get_random_transpose_values(self):
        "per-item semitone shift: randint over transpose_range centred on zero, zeroed with probability 1 - transpose_p (:1057-1063)"
        if self.transpose_range is None:
            return None
        lo, hi = self.transpose_range
        shift = torch.randint(lo, hi, (self.n_items,)) - hi // 2          # torch RNG call 1 (same order as the reference)
        keep = torch.rand(shift.shape) <= self.transpose_p                # torch RNG call 2
        return torch.where(keep, shift, torch.zeros_like(shift))

    def on_epoch_begin(self, **kwargs):
        """New epoch: reshuffle / redraw when `shuffle`, then place the bs row cursors at equal token distances along the epoch's
        item stream: row i starts at token floor(i * totalToks / bs) of the concatenated stream - a searchsorted over the
        cumulative lengths instead of the reference's nested loop (:1065-1084)."""
        if self.idx is None:
            self.allocate_buffers()
        elif self.shuffle:
            self.ite_len = None
            self.idx.shuffle()
            self.transpose_values = self.get_random_transpose_values()
            self.bptt_len = self.bptt
        self.idx.forward = not self.backwards
        len(self)
        lens = self.lengths[self.idx.visit_order()]
        ends = np.cumsum(lens)                                             # ends[k] = tokens in the first k+1 visited items
        starts_at = (self.totalToks / self.bs * np.arange(self.bs)).astype(np.int64)     # int(step * i), step a Python float
        item = np.searchsorted(ends, starts_at, side='right')              # first visited item whose end lies beyond the start
        if self.totalToks == 0:
            item[:] = -1
        into = starts_at - (ends[item] - lens[item])                       # tokens already consumed inside that item
        self.ro[:] = item
        self.ri[:] = lens[item] - into if self.backwards else into
        self._epoch_uploaded = False

    def on_epoch_end(self, **kwargs):                                                # :1087
        self.on_epoch_begin()

    # ---- device side ---------------------------------------------------------------------------------------------
    def _ensure_device(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError('MusicPreloader: no CUDA device - the batches are produced by the CUDA library, there is no CPU path')
            self.device = torch.device('cuda', torch.cuda.current_device())
        if self._dev is None:
            dev = self.device
            off = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
            self._dev = {
                'tokens': torch.from_numpy(self._host_tokens.astype(np.int32)).to(dev),
                'positions': torch.from_numpy(self._host_positions.astype(np.int32)).to(dev) if self._host_positions is not None else None,
                'offsets': torch.from_numpy(off).to(dev),
            }
        if not getattr(self, '_epoch_uploaded', False):
            dev = self.device
            lo, hi = self.rank * self.local_bs, (self.rank + 1) * self.local_bs
            self._dev['perm'] = torch.from_numpy(self.idx.idx.astype(np.int64)).to(dev)
            self._dev['transpose'] = self.transpose_values.to(torch.int32).to(dev) if self.transpose_values is not None else None
            self._dev['ro'] = torch.from_numpy(self.ro[lo:hi].copy()).to(dev)
            self._dev['ri'] = torch.from_numpy(self.ri[lo:hi].copy()).to(dev)
            self._epoch_uploaded = True

    def next_batch(self):
        "the next batch of the running epoch: (x, y) or ({'x': x, 'pos': pos}, y), int64 [bs, bptt] on the device"
        if self.idx is None or not hasattr(self, '_epoch_uploaded'):
            self.on_epoch_begin()
        self._ensure_device()
        d, dev = self._dev, self.device
        x = torch.empty(self.local_bs, self.bptt, dtype=torch.int64, device=dev)
        y = torch.empty_like(x)
        pos = torch.empty_like(x) if self.encode_position else None
        lib = _lib.load()
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.dmg_preload_fill(_p(d['tokens']), _p(d['positions']), _p(d['offsets']), _p(d['perm']), self.n_items,
                                        0 if self.backwards else 1, _p(d['transpose']), self.note_range[0], self.note_range[1],
                                        _p(d['ro']), _p(d['ri']), self.local_bs, self.bptt, self.y_offset, _p(x), _p(y), _p(pos), st),
                   'dmg_preload_fill')
        return ({'x': x, 'pos': pos}, y) if self.encode_position else (x, y)

    def __iter__(self):
        self.on_epoch_begin()
        for _ in range(self.n_batches):
            yield self.next_batch()


def mask_tfm(b, mask_range, mask_idx, pad_idx, p=0.3, seed=None, return_draws=False):
    """deep_music_remix.py:1208-1223 on device tensors: returns the masked copy (x, y) of the batch ``b = (x, y)``.
    The uniform draws and the replacement tokens come from a counter-based generator keyed by ``seed`` (drawn from torch's
    CPU generator when None, so ``torch.manual_seed`` makes the transform reproducible); ``return_draws`` also returns them."""
    x, y = b
    assert x.is_cuda and x.dtype == torch.int64 and y.dtype == torch.int64 and x.shape == y.shape
    x, y = x.clone().contiguous(), y.clone().contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    rand = torch.empty(x.shape, dtype=torch.float32, device=x.device) if return_draws else None
    wrong = torch.empty_like(x) if return_draws else None
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(lib.dmg_mask_tfm(_p(x), _p(y), x.numel(), int(mask_range[0]), int(mask_range[1]), int(mask_idx), int(pad_idx), float(p),
                                seed & 0xFFFFFFFF, _p(rand), _p(wrong), st), 'dmg_mask_tfm')
    return (x, y, rand, wrong) if return_draws else (x, y)
