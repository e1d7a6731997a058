"""Host-side token codec: MIDI file <-> npenc <-> idxenc, vocabulary, beat positions.

CPU-side host logic of the drop-in (not a kernel).  Same names and behaviour as the reference's
``core/vocab.py``, ``core/encodings.py``, ``core/primitives.py`` (= the copies inside
``deep_music_genre.py:126-196, 220-387, 812-890, 1152-1567``), written event-based: notes are reduced per
(time step, part, pitch) directly instead of going through the reference's dense
``[time, part, 128]`` score array.  music21 is replaced by a small Standard-MIDI-File reader/writer.

Bit-exactness is checked in ``tests/test_codec.py`` against the reference's executed-notebook golden
(623 tokens, ``notebooks/Transformer_Genre_Evaluation.ipynb:3299``) and against the oracle restatement.
"""
from enum import Enum
from fractions import Fraction
from functools import partial
import struct

import numpy as np
import torch

BPB = 4
TIMESIG = f'{BPB}/4'
SAMPLE_FREQ = 4
NOTE_SIZE = 128
DUR_SIZE = (10 * BPB * SAMPLE_FREQ) + 1
MAX_NOTE_DUR = 8 * BPB * SAMPLE_FREQ
NOTE_RANGE = (1, 127)
VALTSEP, VALTCONT = -1, -2

BOS, PAD, EOS, MASK, SEP, IN = 'xxbos', 'xxpad', 'xxeos', 'xxmask', 'xxsep', 'xxni'
ELECTRONIC, FOLK, FUNK, JAZZ, POP, ROCK = 'xxelec', 'xxfolk', 'xxfunk', 'xxjazz', 'xxpop', 'xxrock'
GENRE_TOKS = {'electronic': ELECTRONIC, 'folk': FOLK, 'funk': FUNK, 'jazz': JAZZ, 'pop': POP, 'rock': ROCK}
ACCEP_INS = {'Piano': 0, 'Guitar': 1, 'Bass': 2, 'WoodwindInstrument': 3, 'BrassInstrument': 4, 'StringInstrument': 5,
             'Misc': 6}
ACCEP_INS_REV = {v: k for k, v in ACCEP_INS.items()}
NOTE_TOKS = [f'n{i}' for i in range(NOTE_SIZE)]
DUR_TOKS = [f'd{i}' for i in range(DUR_SIZE)]
INS_TOKS = [f'i{i}' for i in range(len(ACCEP_INS))]
MTEMPO_SIZE = 10
MTEMPO_TOKS = [f'mt{i}' for i in range(MTEMPO_SIZE)]
SPECIAL_TOKS = [BOS, PAD, EOS, MASK, ELECTRONIC, FOLK, FUNK, JAZZ, POP, ROCK, IN, SEP]
NULL_INS = -2 - len(NOTE_TOKS) - len(DUR_TOKS)        # instrument column of separator rows

SEQType = Enum('SEQType', 'Mask, Sentence, Melody, Chords, Empty, Genre')


class MusicVocab:
    "Token <-> index mapping (reference: deep_music_genre.py:812-890 / core/vocab.py)."
    def __init__(self, itos):
        self.itos = list(itos)
        self.stoi = {tok: i for i, tok in enumerate(self.itos)}

    @classmethod
    def create(cls):
        itos = SPECIAL_TOKS + NOTE_TOKS + DUR_TOKS + INS_TOKS + MTEMPO_TOKS
        rem = len(itos) % 8
        if rem:                       # the reference pads with len%8 dummies (318 -> 324), not up to a multiple of 8
            itos = itos + [f'dummy{i}' for i in range(rem)]
        return cls(itos)

    def numericalize(self, toks): return [self.stoi[t] for t in toks]

    def textify(self, nums, sep=' '):
        toks = [self.itos[int(i)] for i in nums]
        return sep.join(toks) if sep is not None else toks

    def to_music_item(self, idxenc, ins=None): return MusicItem(idxenc, self, ins)

    mask_idx = property(lambda s: s.stoi[MASK])
    pad_idx = property(lambda s: s.stoi[PAD])
    bos_idx = property(lambda s: s.stoi[BOS])
    sep_idx = property(lambda s: s.stoi[SEP])
    ni_idx = property(lambda s: s.stoi[IN])
    npenc_range = property(lambda s: (s.stoi[IN], s.stoi[INS_TOKS[-1]] + 1))
    note_range = property(lambda s: (s.stoi[NOTE_TOKS[0]], s.stoi[NOTE_TOKS[-1]] + 1))
    dur_range = property(lambda s: (s.stoi[DUR_TOKS[0]], s.stoi[DUR_TOKS[-1]] + 1))
    ins_range = property(lambda s: (s.stoi[INS_TOKS[0]], s.stoi[INS_TOKS[-1]] + 1))

    def is_duration(self, idx): return self.dur_range[0] <= idx < self.dur_range[1]
    def is_duration_or_pad(self, idx): return idx == self.pad_idx or self.is_duration(idx)
    def is_note(self, idx): return idx == self.sep_idx or self.note_range[0] <= idx < self.note_range[1]
    def is_ins(self, idx): return idx == self.ni_idx or self.ins_range[0] <= idx < self.ins_range[1]
    def __len__(self): return len(self.itos)
    def __getstate__(self): return {'itos': self.itos}

    def __setstate__(self, state):
        self.itos = state['itos']
        self.stoi = {tok: i for i, tok in enumerate(self.itos)}


# ------------------------------------------------------------------------------------------ MIDI file reader
def _varlen(buf, i):
    val = 0
    while True:
        byte = buf[i]
        i += 1
        val = (val << 7) | (byte & 0x7F)
        if byte < 0x80:
            return val, i


def parse_midi(path):
    """Standard MIDI File -> (ticks_per_quarter, [track]), track = {'notes': [(on_tick, pitch, ticks)], 'programs':
    [(tick, program)]}.  Note-on with velocity 0 counts as note-off; on/off are paired first-in-first-out per
    (channel, pitch); running status is honoured."""
    with open(path, 'rb') as fh:
        raw = fh.read()
    if raw[:4] != b'MThd':
        raise ValueError(f'{path}: not a Standard MIDI File')
    hdr_len, = struct.unpack('>I', raw[4:8])
    _fmt, ntracks, division = struct.unpack('>HHH', raw[8:14])
    if division & 0x8000:
        raise ValueError('SMPTE time division is not supported')
    at = 8 + hdr_len
    tracks = []
    for _ in range(ntracks):
        if raw[at:at + 4] != b'MTrk':
            raise ValueError('bad track chunk')
        size, = struct.unpack('>I', raw[at + 4:at + 8])
        body = raw[at + 8:at + 8 + size]
        at += 8 + size
        i, now, status = 0, 0, None
        pending, notes, programs = {}, [], []
        while i < len(body):
            delta, i = _varlen(body, i)
            now += delta
            if body[i] & 0x80:
                status = body[i]
                i += 1
            if status == 0xFF:                       # meta event
                i += 1
                n, i = _varlen(body, i)
                i += n
            elif status in (0xF0, 0xF7):             # sysex
                n, i = _varlen(body, i)
                i += n
            else:
                kind, ch = status & 0xF0, status & 0x0F
                if kind in (0xC0, 0xD0):
                    if kind == 0xC0:
                        programs.append((now, body[i]))
                    i += 1
                else:
                    d1, d2 = body[i], body[i + 1]
                    i += 2
                    if kind == 0x90 and d2 > 0:
                        pending.setdefault((ch, d1), []).append(now)
                    elif kind == 0x80 or kind == 0x90:
                        starts = pending.get((ch, d1))
                        if starts:
                            t0 = starts.pop(0)
                            notes.append((t0, d1, now - t0))
        tracks.append({'notes': notes, 'programs': programs})
    return division, tracks


def _snap(q):
    "music21's default quantiser: nearest point of the 1/4- or the 1/3-quarter grid (first best wins)."
    a, b = Fraction(round(q * 4), 4), Fraction(round(q * 3), 3)
    return a if abs(a - q) <= abs(b - q) else b


def _program_class(prog):
    """General-MIDI program -> the ACCEP_INS class the reference's instrument test (deep_music_genre.py:251-291)
    reaches through music21's instrument classes; None = instrument rejected."""
    if prog <= 8 or prog == 55 or 80 <= prog <= 103 or prog >= 117: return 'Piano'
    if prog == 15 or prog == 32 or 40 <= prog <= 46 or 48 <= prog <= 51 or 104 <= prog <= 107 or prog == 110:
        return 'StringInstrument'
    if 24 <= prog <= 31: return 'Guitar'
    if 33 <= prog <= 39: return 'Bass'
    if 56 <= prog <= 63: return 'BrassInstrument'
    if 64 <= prog <= 79 or prog in (109, 111): return 'WoodwindInstrument'
    return None


def midi2npenc(path, sample_freq=SAMPLE_FREQ, max_note_dur=MAX_NOTE_DUR, note_range=NOTE_RANGE):
    """MIDI file -> (npenc [n, 3] rows [note, dur, part] with separator rows [-1, wait, NULL_INS], ins {part: class}).
    Equivalent to file2stream -> stream2chordarr -> chordarr2npenc of the reference."""
    tpq, tracks = parse_midi(path)
    parts = [t for t in tracks if t['notes']]
    ins = {}
    cells = {}                      # (step, part, pitch) -> duration of the onset that survives in the score grid
    for part, trk in enumerate(parts):
        # walk the part in time order: instruments first at equal offsets, like music21's flattened stream
        events = [(_snap(Fraction(t, tpq)), 0, ('ins', p)) for t, p in trk['programs']]
        events += [(_snap(Fraction(t, tpq)), 1, ('note', pitch, _snap(Fraction(n, tpq)))) for t, pitch, n in trk['notes']]
        events.sort(key=lambda e: (e[0], e[1]))
        accepted, notes = False, []
        for off, _, ev in events:
            if ev[0] == 'ins':
                cls_ = _program_class(ev[1])
                if cls_ is None:
                    break                              # rejected instrument ends the scan of this part
                ins[part], accepted = cls_, True
            else:
                notes.append((int(round(off * sample_freq)), int(round(ev[2] * sample_freq)), ev[1]))
        if not accepted:
            continue
        for step, dur, pitch in sorted(notes, key=lambda n: (n[0], n[1])):   # later (longer) notes overwrite
            cells[(step, part, pitch)] = min(dur, max_note_dur) if max_note_dur is not None else dur
    by_step = {}
    for (step, part, pitch), dur in cells.items():
        if dur > 0 and note_range[0] <= pitch < note_range[1]:
            by_step.setdefault(step, []).append((pitch, dur, part))
    rows, last = [], None
    for step in sorted(by_step):
        wait = step if last is None else step - last
        if wait > 0:
            rows.append((VALTSEP, wait, NULL_INS))
        rows.extend(sorted(by_step[step], key=lambda r: (-r[0], r[2])))    # pitch high->low, then part
        last = step
    return np.array(rows, dtype=int).reshape(-1, 3), ins


# ------------------------------------------------------------------------------------------ npenc <-> idxenc
def sort_instruments(npenc, vocab=None):
    """Within every separator-delimited group order rows by part index (stable).  Keeps the reference's handling of
    the trailing group, which re-emits the second-to-last separator row (deep_music_genre.py:1469-1473)."""
    seps = np.flatnonzero(npenc[:, 0] == VALTSEP)
    if len(seps) == 0:
        raise IndexError('sort_instruments: no separator row')     # reference: sep_idxs[0] on an empty array
    def ordered(block): return block[np.argsort(block[:, 2], kind='stable')] if len(block) else block
    out = []
    if seps[0] != 0:
        out.append(ordered(npenc[:seps[0]]))
    for a, b in zip(seps[:-1], seps[1:]):
        out.append(npenc[a:a + 1])
        out.append(ordered(npenc[a + 1:b]))
    if len(seps) < 2:
        raise NameError('sort_instruments: a single separator row (the reference fails here too)')
    tail_sep = npenc[seps[-2]:seps[-2] + 1]
    out.append(tail_sep)
    if len(npenc) > seps[-1] + 1:
        out.append(ordered(npenc[seps[-1] + 1:]))
    res = np.concatenate(out, axis=0)
    assert list(seps) == list(np.flatnonzero(res[:, 0] == VALTSEP))
    return res


def npins2vocabins(x, ins):
    if x in ins:
        return ACCEP_INS.get(ins[x], ACCEP_INS['Piano'])
    if x == NULL_INS:
        return x
    raise Exception(f'unknown part index {x}')


def seq_prefix(seq_type, vocab, genre=None):
    if seq_type == SEQType.Empty:
        return np.empty(0, dtype=int)
    start = vocab.bos_idx
    if seq_type == SEQType.Genre and genre is not None:
        g = genre.lower()
        for key, tok in GENRE_TOKS.items():
            if key in g:
                start = vocab.stoi[tok]
                break
    return np.array([start, vocab.pad_idx])


def npenc2idxenc(t, vocab, ins=None, genre=None, seq_type=SEQType.Sentence, add_eos=True):
    "[[n, d, i], ...] -> flat ids with the [bos|genre, pad] prefix and the eos suffix."
    t = np.array(t, dtype=int).copy()
    t[:, 0] += vocab.note_range[0]
    t[:, 1] += vocab.dur_range[0]
    if t.shape[1] == 3:
        if ins is not None:
            t[:, 2] = [npins2vocabins(x, ins) for x in t[:, 2]]
        t[:, 2] += vocab.ins_range[0]
        prefix = seq_prefix(seq_type, vocab, genre)
    else:
        prefix = seq_prefix(seq_type, vocab)
    suffix = np.array([vocab.stoi[EOS]]) if add_eos else np.empty(0, dtype=int)
    return np.concatenate([prefix, t.reshape(-1), suffix])


def to_valid_idxenc(t, valid_range):
    return t[(t >= valid_range[0]) & (t < valid_range[1])]


def to_valid_npenc(t):
    bad_note = (t[:, 0] < VALTSEP) | (t[:, 0] >= NOTE_SIZE)
    i_note, i_dur = bad_note.argmax(), (t[:, 1] < 0).argmax()
    cut = max(i_dur, i_note)
    if cut > 0:
        if i_note > 0 and i_dur > 0:
            cut = min(i_dur, i_note)
        print('Non midi note detected. Only returning valid portion. Index, seed', cut, t.shape)
        return t[:cut]
    return t


def idxenc2npenc(t, vocab, validate=True):
    t = np.asarray(t)
    if validate:
        t = to_valid_idxenc(t, vocab.npenc_range)
    is_ins = [bool(vocab.is_ins(x)) for x in t]
    t = t[:len(is_ins) - is_ins[::-1].index(True)]          # cut after the last instrument token
    t = t.copy().reshape(-1, 3)
    if t.shape[0] == 0:
        return t
    t[:, 0] -= vocab.note_range[0]
    t[:, 1] -= vocab.dur_range[0]
    t[:, 2] -= vocab.ins_range[0]
    return to_valid_npenc(t) if validate else t


def position_enc(idxenc, vocab):
    "Beat position of every token: cumulative sum of the separator durations, placed 3 tokens after each xxsep."
    idxenc = np.asarray(idxenc)
    seps = np.flatnonzero(idxenc == vocab.sep_idx)
    seps = seps[seps + 2 < idxenc.shape[0]]
    durs = idxenc[seps + 1].copy()
    durs[durs == vocab.mask_idx] = vocab.dur_range[0]
    durs -= vocab.dur_range[0]
    pos = np.zeros_like(idxenc)
    if len(seps):
        if len(idxenc) <= seps[-1] + 3:
            seps, durs = seps[:-1], durs[:-1]
        pos[seps + 3] = durs
    return pos.cumsum()


def find_beat(pos, beat, sample_freq=SAMPLE_FREQ, side='left'):
    return np.searchsorted(pos, beat * sample_freq, side=side)


def beat2index(idxenc, pos, vocab, beat, include_last_sep=False):
    cutoff = find_beat(pos, beat)
    if cutoff < 2:
        return 2
    if len(idxenc) < 2 or include_last_sep:
        return cutoff
    return cutoff - 2 if idxenc[cutoff - 2] == vocab.sep_idx else cutoff


def trim_to_beat(idxenc, pos, vocab, to_beat=None, include_last_sep=True):
    if to_beat is None:
        return idxenc
    return idxenc[:beat2index(idxenc, pos, vocab, to_beat, include_last_sep=include_last_sep)]


def tfm_transpose(x, value, vocab):
    x = x.copy()
    x[(x >= vocab.note_range[0]) & (x < vocab.note_range[1])] += value
    return x


def mask_section(xb, pos, token_range, replacement_idx, section_range=None):
    xb = xb.copy()
    tok = (xb >= token_range[0]) & (xb < token_range[1])
    lo, hi = section_range if section_range is not None else (None, None)
    a = find_beat(pos, lo) if lo is not None else 0
    b = find_beat(pos, hi) if hi is not None else xb.shape[0]
    sec = np.zeros_like(xb, dtype=bool)
    sec[a:b] = True
    xb[tok & sec] = replacement_idx
    return xb


def pad_seq(seq, bptt, value):
    return np.pad(seq, (0, max(bptt - seq.shape[0], 0)), 'constant', constant_values=value)[:bptt]


def to_tensor(t, device=None):
    t = t if isinstance(t, torch.Tensor) else torch.tensor(np.asarray(t))
    if device is None and torch.cuda.is_available():
        t = t.cuda()
    elif device is not None:
        t = t.to(device)
    return t.long()


# ------------------------------------------------------------------------------------------ MIDI writer
class MidiStream:
    """What ``MusicItem.to_stream(bpm)`` returns: enough of music21's Stream for ``.write('midi', fp=...)``
    (app.py:191, 275).  One tempo track plus one track per instrument class, 1024 ticks per quarter (layout: ``to_bytes``)."""
    TPQ = 1024
    PROGRAMS = {0: 0, 1: 24, 2: 33, 3: 73, 4: 61, 5: 40, 6: 0}      # ACCEP_INS index -> General-MIDI program

    def __init__(self, npenc, bpm=120):
        self.npenc, self.bpm = np.asarray(npenc), bpm

    def notes(self):
        "[(start_step, dur_steps, pitch, ins)]"
        out, now = [], 0
        for row in self.npenc:
            n, d = int(row[0]), int(row[1])
            i = int(row[2]) if len(row) > 2 else 0
            if n == VALTSEP:
                now += d
            elif n >= 0:
                out.append((now, d, n, max(i, 0)))
        return out

    @staticmethod
    def _vl(n):
        out = [n & 0x7F]
        n >>= 7
        while n:
            out.append((n & 0x7F) | 0x80)
            n >>= 7
        return bytes(reversed(out))

    NAMES = {0: 'Piano', 1: 'Guitar', 2: 'Bass', 3: 'Woodwind', 4: 'Brass', 5: 'StringInstrument', 6: 'Piano'}

    def _track(self, events, end_gap):
        "events: (tick, order, bytes); same-tick events keep `order` (note-offs before note-ons: offs by note start then pitch, ons by duration group then pitch); EOT `end_gap` ticks later"
        body, now = b'', 0
        for tick, _, data in sorted(events, key=lambda e: (e[0], e[1])):
            body += self._vl(tick - now) + data
            now = tick
        body += self._vl(end_gap) + b'\xff\x2f\x00'
        return b'MTrk' + struct.pack('>I', len(body)) + body

    def to_bytes(self):
        """The file layout music21 gives ``full.to_stream(bpm).write('midi', fp)`` (app.py:191, 275; pinned byte for byte by the
        reference's own outputs/genre_output.mid and outputs/remix_Notes_output.mid for the Piano-only case): format 1, 1024 ticks
        per quarter; track 0 = tempo, key signature 0, time signature 4/4, end of track one quarter later; one track per part =
        track name, program change, pitch-bend centre, program change again, then note-on velocity 90 / note-off velocity 0 at the
        exact note end, the end of track one quarter after the last event."""
        per = self.TPQ // SAMPLE_FREQ
        tempo = int(round(60_000_000 / self.bpm))
        conductor = [(0, 0, b'\xff\x51\x03' + tempo.to_bytes(3, 'big')), (0, 1, b'\xff\x59\x02\x00\x00'),
                     (0, 2, b'\xff\x58\x04\x04\x02\x18\x08')]
        tracks = [self._track(conductor, self.TPQ)]
        by_ins = {}
        for start, dur, pitch, ins in self.notes():
            by_ins.setdefault(ins, []).append((start, dur, pitch))
        for ch, (ins, notes) in enumerate(sorted(by_ins.items())):
            ch = min(ch if ch < 9 else ch + 1, 15)                 # channel 10 (index 9) is percussion
            name = self.NAMES.get(ins, 'Piano').encode()
            prog = bytes([0xC0 | ch, self.PROGRAMS.get(ins, 0)])
            ev = [(0, (-4,), b'\xff\x03' + self._vl(len(name)) + name), (0, (-3,), prog), (0, (-2,), bytes([0xE0 | ch, 0x00, 0x40])), (0, (-1,), prog)]
            for start, dur, pitch in notes:
                ev.append((start * per, (1, dur, pitch), bytes([0x90 | ch, pitch, 90])))     # group_notes_by_duration (:536-541)
                ev.append(((start + dur) * per, (0, start, pitch), bytes([0x80 | ch, pitch, 0])))
            tracks.append(self._track(ev, self.TPQ))
        return b'MThd' + struct.pack('>IHHH', 6, 1, len(tracks), self.TPQ) + b''.join(tracks)

    def write(self, fmt='midi', fp=None):
        assert fmt in ('midi', 'mid')
        with open(fp, 'wb') as fh:
            fh.write(self.to_bytes())
        return fp


# ------------------------------------------------------------------------------------------ MusicItem
class MusicItem:
    "Reference: deep_music_genre.py:1152-1278 / core/primitives.py:10-137."
    def __init__(self, data, vocab, ins=None, verbose=False, stream=None, position=None):
        self.data, self.vocab, self.ins = data, vocab, ins
        self._stream, self._position = stream, position

    def __repr__(self):
        return '\n'.join([f'\n{self.__class__.__name__} - {self.data.shape}', f'npenc: {self.data[:10]}',
                          f'{self.vocab.textify(self.data[:10])}...'])

    def __len__(self): return len(self.data)

    @classmethod
    def from_file(cls, midi_file, vocab):
        npenc, ins = midi2npenc(midi_file)
        cls.ins = ins                                    # the reference sets the class attribute too (:1172)
        return cls.from_npenc(npenc, vocab, None, ins)

    @classmethod
    def from_npenc(cls, npenc, vocab, stream=None, ins=None, genre=None):
        npenc = sort_instruments(npenc, vocab)
        seq_type = SEQType.Genre if genre is not None else SEQType.Sentence
        return MusicItem(npenc2idxenc(npenc, vocab, ins=ins, genre=genre, seq_type=seq_type), vocab, ins=ins, stream=stream)

    @classmethod
    def from_idx(cls, item, vocab):
        idx, pos = item
        return MusicItem(idx, vocab=vocab, position=pos)

    def to_idx(self): return self.data, self.position

    @classmethod
    def empty(cls, vocab, seq_type=SEQType.Sentence): return MusicItem(seq_prefix(seq_type, vocab), vocab)

    @property
    def stream(self):
        if self._stream is None:
            self._stream = self.to_stream()
        return self._stream

    def to_stream(self, bpm=120): return MidiStream(idxenc2npenc(self.data, self.vocab), bpm=bpm)
    def to_tensor(self, device=None): return to_tensor(self.data, device)
    def to_text(self, sep=' '): return self.vocab.textify(self.data, sep)

    @property
    def position(self):
        if self._position is None:
            self._position = position_enc(self.data, self.vocab)
        return self._position

    def get_pos_tensor(self, device=None): return to_tensor(self.position, device)
    def to_npenc(self): return idxenc2npenc(self.data, self.vocab)

    @property
    def new(self): return partial(type(self), vocab=self.vocab)

    def trim_to_beat(self, beat, include_last_sep=False):
        return self.new(trim_to_beat(self.data, self.position, self.vocab, beat, include_last_sep))

    def transpose(self, interval):
        return self.new(tfm_transpose(self.data, interval, self.vocab), position=self._position)

    def append(self, item): return self.new(np.concatenate((self.data, item.data), axis=0))
    def mask_pitch(self, section=None): return self.new(self.mask(self.vocab.note_range, section), position=self.position)

    def mask_duration(self, section=None, keep_position_enc=True):
        masked = self.mask(self.vocab.dur_range, section)
        return self.new(masked, position=self.position) if keep_position_enc else self.new(masked)

    def mask(self, token_range, section_range=None):
        return mask_section(self.data, self.position, token_range, self.vocab.mask_idx, section_range=section_range)

    def pad_to(self, bptt):
        return self.new(pad_seq(self.data, bptt, self.vocab.pad_idx), stream=self._stream,
                        position=pad_seq(self.position, bptt, 0))

    def remove_eos(self):
        return self.new(self.data, stream=self.stream) if self.data[-1] == self.vocab.stoi[EOS] else self


class MusicDataBunch:
    "Only what the inference call sites use: ``MusicDataBunch.empty(path).vocab`` (app_utils.py:74, 80)."
    def __init__(self, vocab, path=''):
        self.vocab, self.path = vocab, path

    @classmethod
    def empty(cls, path='', **kwargs): return cls(MusicVocab.create(), path)
