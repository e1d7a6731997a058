"""TEST INFRASTRUCTURE - CPU restatement of the reference's training data feed.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(deepmusicgeneration_b200/) never does.

Restates, literally and in numpy:
  * ``MusicPreloader`` (deep_music_genre.py:1001-1125): the fastai ``LanguageModelPreLoader`` variant that turns a ragged list of
    token arrays into ``bs`` contiguous streams, ``bptt`` tokens per batch, y = x shifted by ``y_offset``, with a per-item random
    transpose (``MusicItem.transpose`` :1247 -> ``tfm_transpose`` :1541-1544) and, with ``encode_position``, the beat positions
    stacked behind the indices (``batch_position_tfm`` :1129-1136);
  * ``mask_tfm`` (deep_music_remix.py:1208-1223): the BERT-style masking of a batch for the remix encoder.
Pinned against the reference's own code: tests/golden/make_preloader_golden.py executes the reference class source (with the
fastai base class stubbed) and stores its batches in tests/golden/preloader_golden.npz; tests/test_preloader_cpu.py compares.
"""
import math

import numpy as np
import torch


class Item:
    "the two arrays of a MusicItem the preloader touches (deep_music_genre.py:1160-1249): .data (token ids) and .position"

    def __init__(self, data, position):
        self.data, self.position = np.asarray(data, dtype=np.int64), np.asarray(position, dtype=np.int64)

    def __len__(self):
        return len(self.data)


def tfm_transpose(x, value, note_range):
    "deep_music_genre.py:1541-1544"
    x = x.copy()
    x[(x >= note_range[0]) & (x < note_range[1])] += value
    return x


class CircularIndex:
    "deep_music_genre.py:1005-1014"

    def __init__(self, length, forward):
        self.idx, self.forward = np.arange(length), forward

    def __getitem__(self, i):
        return self.idx[i % len(self.idx) if self.forward else len(self.idx) - 1 - i % len(self.idx)]

    def __len__(self):
        return len(self.idx)

    def shuffle(self):
        np.random.shuffle(self.idx)


class MusicPreloader:
    "deep_music_genre.py:1001-1125 (world size 1: ``bs *= num_distrib() or 1`` is the identity)"

    def __init__(self, items, note_range, bs=32, bptt=70, backwards=False, shuffle=False, y_offset=1, transpose_range=None,
                 transpose_p=0.5, encode_position=True):
        self.items, self.note_range = list(items), tuple(note_range)
        self.bs, self.bptt, self.shuffle, self.backwards = bs, bptt, shuffle, backwards
        self.lengths = None
        self.totalToks, self.ite_len, self.idx = 0, None, None
        self.y_offset = y_offset
        self.transpose_range, self.transpose_p = transpose_range, transpose_p
        self.encode_position = encode_position
        self.bptt_len = self.bptt
        self.allocate_buffers()

    def __len__(self):                                                               # :1032-1037
        if self.ite_len is None:
            if self.lengths is None:
                self.lengths = np.array([len(item) for item in self.items])
            self.totalToks = self.lengths.sum()
            self.ite_len = self.bs * int(math.ceil(self.totalToks / (self.bptt * self.bs)))
        return self.ite_len

    def allocate_buffers(self):                                                      # :1041-1055
        if self.ite_len is None:
            len(self)
        self.idx = CircularIndex(len(self.items), not self.backwards)
        buffer_len = (2,) if self.encode_position else ()
        self.batch = np.zeros((self.bs, self.bptt + self.y_offset) + buffer_len, dtype=np.int64)
        self.batch_x, self.batch_y = self.batch[:, 0:self.bptt], self.batch[:, self.y_offset:self.bptt + self.y_offset]
        self.ro = np.zeros(self.bs, dtype=np.int64)
        self.ri = np.zeros(self.bs, dtype=np.int64)
        self.transpose_values = self.get_random_transpose_values()

    def get_random_transpose_values(self):                                           # :1057-1063
        if self.transpose_range is None:
            return None
        n = len(self.items)
        rt_arr = torch.randint(*self.transpose_range, (n,)) - self.transpose_range[1] // 2
        mask = torch.rand(rt_arr.shape) > self.transpose_p
        rt_arr[mask] = 0
        return rt_arr

    def on_epoch_begin(self):                                                        # :1065-1084
        if self.idx is None:
            self.allocate_buffers()
        elif self.shuffle:
            self.ite_len = None
            self.idx.shuffle()
            self.transpose_values = self.get_random_transpose_values()
            self.bptt_len = self.bptt
        self.idx.forward = not self.backwards
        step = self.totalToks / self.bs
        ln_rag, countTokens, i_rag = 0, 0, -1
        for i in range(0, self.bs):
            while ln_rag + countTokens <= int(step * i):
                countTokens += ln_rag
                i_rag += 1
                ln_rag = self.lengths[self.idx[i_rag]]
            self.ro[i] = i_rag
            self.ri[i] = (ln_rag - int(step * i - countTokens)) if self.backwards else int(step * i - countTokens)

    def __getitem__(self, k):                                                        # :1088-1096
        j = k % self.bs
        self.ro[j], self.ri[j] = self.fill_row(not self.backwards, self.items, self.idx, self.batch[j][:self.bptt_len + self.y_offset],
                                               self.ro[j], self.ri[j], overlap=1, lengths=self.lengths)
        return self.batch_x[j][:self.bptt_len], self.batch_y[j][:self.bptt_len]

    def fill_row(self, forward, items, idx, row, ro, ri, overlap, lengths):          # :1098-1125
        ibuf = n = 0
        ro -= 1
        while ibuf < row.shape[0]:
            ro += 1
            ix = idx[ro]
            item = items[ix]
            data = item.data
            if self.transpose_values is not None:
                data = tfm_transpose(data, self.transpose_values[ix].item(), self.note_range)
            rag = np.stack([data, item.position], axis=1) if self.encode_position else data
            if forward:
                ri = 0 if ibuf else ri
                n = min(lengths[ix] - ri, row.shape[0] - ibuf)
                row[ibuf:ibuf + n] = rag[ri:ri + n]
            else:
                ri = lengths[ix] if ibuf else ri
                n = min(ri, row.size - ibuf)          # sic (:1120): row.size, not row.shape[0] - with encode_position a backwards
                                                      # epoch therefore fails in the reference as soon as n exceeds the rows left
                row[ibuf:ibuf + n] = rag[ri - n:ri][::-1]
            ibuf += n
        return ro, ri + ((n - overlap) if forward else -(n - overlap))

    def batches(self):
        "one epoch the way fastai's DataLoader consumes the preloader: items k = 0 .. len-1 in order, bs rows per batch"
        self.on_epoch_begin()
        for b in range(len(self) // self.bs):
            xs, ys = [], []
            for j in range(self.bs):
                x, y = self[b * self.bs + j]
                xs.append(x.copy()); ys.append(y.copy())
            x, y = np.stack(xs), np.stack(ys)
            if self.encode_position:                                                  # batch_position_tfm, :1129-1136
                yield {'x': x[..., 0], 'pos': x[..., 1]}, y[..., 0]
            else:
                yield x, y


def mask_tfm(x, y, mask_range, mask_idx, pad_idx, p=0.3, rand=None, wrong=None):
    """deep_music_remix.py:1208-1223 on torch tensors.  ``rand`` (uniform [0,1), x.shape) and ``wrong`` (the replacement tokens of
    the 10 % 'wrong word' positions, in row-major order of those positions) may be injected so that a device implementation
    that draws them itself can be compared element for element."""
    x, y = x.clone(), y.clone()
    rand = torch.rand(x.shape, device=x.device) if rand is None else rand.clone()
    rand[x < mask_range[0]] = 1.0
    rand[x >= mask_range[1]] = 1.0
    y[rand > p] = pad_idx
    x[rand <= (p * .8)] = mask_idx
    wrong_word = (rand > (p * .8)) & (rand <= (p * .9))
    n = int(wrong_word.sum().item())
    x[wrong_word] = torch.randint(*mask_range, [n], device=x.device) if wrong is None else wrong[:n].to(x.dtype)
    return x, y
