"""Oracle: vocab + MIDI -> npenc -> idxenc codec.  TEST INFRASTRUCTURE.

A literal (dense-array, loop-by-loop) restatement of ``deep_music_genre.py``: constants ``:126-196``,
``stream2chordarr`` ``:220-322``, ``chordarr2npenc`` ``:324-345``, ``timestep2npenc`` ``:367-387``,
``MusicVocab`` ``:812-890``, ``npins2vocabins``/``npenc2idxenc``/``seq_prefix`` ``:1301-1376``,
``idxenc2npenc``/``to_valid_*`` ``:1380-1441``, ``sort_instruments`` ``:1443-1487``, ``position_enc``
``:1489-1527``, ``beat2index``/``find_beat``/``trim_to_beat`` ``:1529-1549`` (the same code is duplicated in
``core/encodings.py``, ``core/primitives.py``, ``core/vocab.py``).

music21 (``file2stream`` ``:210-212``) is an UN-VENDORED, version-unpinned dependency; it is replaced by the
plain Standard-MIDI-File reader below, which reproduces what music21's MIDI translation hands to
``stream2chordarr``: per note-bearing track one part, per note ``(pitch, offset, quarterLength)`` in quarter
lengths, quantised to the nearer of the 1/4 and 1/3 quarter grids (music21 ``quantize`` default
``quarterLengthDivisors=(4, 3)``), and one Instrument per MIDI program change.

Pinning: the whole pipeline reproduces the reference's 623-token Megalovania known answer
(``notebooks/Transformer_Genre_Evaluation.ipynb:3299``) exactly.  NOT pinned by anything in the reference
(``PARITY UNPINNED``): off-grid quantisation (fur_elise.mid), the GM-program -> instrument-class table
(anything but program 0), multi-instrument tracks.
"""
from fractions import Fraction
import struct

import numpy as np

# ------------------------------------------------------------------ constants (deep_music_genre.py:126-196)
BPB = 4
SAMPLE_FREQ = 4
NOTE_SIZE = 128
DUR_SIZE = (10 * BPB * SAMPLE_FREQ) + 1
MAX_NOTE_DUR = (8 * BPB * SAMPLE_FREQ)
NOTE_RANGE = (1, 127)
VALTSEP = -1
VALTCONT = -2

BOS, PAD, EOS, MASK, SEP, IN = 'xxbos', 'xxpad', 'xxeos', 'xxmask', 'xxsep', 'xxni'
ELECTRONIC, FOLK, FUNK, JAZZ, POP, ROCK = 'xxelec', 'xxfolk', 'xxfunk', 'xxjazz', 'xxpop', 'xxrock'
ACCEP_INS = {'Piano': 0, 'Guitar': 1, 'Bass': 2, 'WoodwindInstrument': 3, 'BrassInstrument': 4,
             'StringInstrument': 5, 'Misc': 6}
NOTE_TOKS = [f'n{i}' for i in range(NOTE_SIZE)]
DUR_TOKS = [f'd{i}' for i in range(DUR_SIZE)]
INS_TOKS = [f'i{i}' for i in range(len(ACCEP_INS))]
MTEMPO_TOKS = [f'mt{i}' for i in range(10)]
SPECIAL_TOKS = [BOS, PAD, EOS, MASK, ELECTRONIC, FOLK, FUNK, JAZZ, POP, ROCK, IN, SEP]
NI_NPENC = -2 - len(NOTE_TOKS) - len(DUR_TOKS)      # -291, the "null instrument" column of separator rows


class MusicVocab:
    "deep_music_genre.py:812-890"
    def __init__(self, itos):
        self.itos = itos
        self.stoi = {v: k for k, v in enumerate(self.itos)}

    def numericalize(self, t): return [self.stoi[w] for w in t]

    def textify(self, nums, sep=' '):
        items = [self.itos[i] for i in nums]
        return sep.join(items) if sep is not None else items

    @property
    def mask_idx(self): return self.stoi[MASK]
    @property
    def pad_idx(self): return self.stoi[PAD]
    @property
    def bos_idx(self): return self.stoi[BOS]
    @property
    def sep_idx(self): return self.stoi[SEP]
    @property
    def ni_idx(self): return self.stoi[IN]
    @property
    def npenc_range(self): return (self.stoi[IN], self.stoi[INS_TOKS[-1]] + 1)
    @property
    def note_range(self): return self.stoi[NOTE_TOKS[0]], self.stoi[NOTE_TOKS[-1]] + 1
    @property
    def dur_range(self): return self.stoi[DUR_TOKS[0]], self.stoi[DUR_TOKS[-1]] + 1
    @property
    def ins_range(self): return self.stoi[INS_TOKS[0]], self.stoi[INS_TOKS[-1]] + 1

    def is_duration(self, idx): return idx >= self.dur_range[0] and idx < self.dur_range[1]
    def is_duration_or_pad(self, idx): return idx == self.pad_idx or self.is_duration(idx)
    def is_note(self, idx): return idx == self.sep_idx or (idx >= self.note_range[0] and idx < self.note_range[1])
    def is_ins(self, idx): return idx == self.ni_idx or (idx >= self.ins_range[0] and idx < self.ins_range[1])
    def __len__(self): return len(self.itos)

    @classmethod
    def create(cls):
        itos = SPECIAL_TOKS + NOTE_TOKS + DUR_TOKS + INS_TOKS + MTEMPO_TOKS
        if len(itos) % 8 != 0:
            itos = itos + [f'dummy{i}' for i in range(len(itos) % 8)]     # sic: len%8, not 8-len%8
        return cls(itos)


# ------------------------------------------------------------------ SMF reader (stands in for music21)
def _vlq(d, i):
    v = 0
    while True:
        c = d[i]; i += 1
        v = (v << 7) | (c & 0x7f)
        if not c & 0x80:
            return v, i


def read_smf(path):
    """-> (ticks_per_quarter, tracks); each track = list of events in file order:
    ('on', tick, ch, pitch) / ('off', tick, ch, pitch) / ('prog', tick, ch, program)."""
    b = open(path, 'rb').read()
    assert b[:4] == b'MThd', 'not a Standard MIDI File'
    hlen = struct.unpack('>I', b[4:8])[0]
    fmt, ntr, div = struct.unpack('>HHH', b[8:14])
    assert not div & 0x8000, 'SMPTE time division not supported'
    p = 8 + hlen
    tracks = []
    for _ in range(ntr):
        assert b[p:p + 4] == b'MTrk'
        ln = struct.unpack('>I', b[p + 4:p + 8])[0]
        d = b[p + 8:p + 8 + ln]; p += 8 + ln
        i, tick, rs, ev = 0, 0, None, []
        while i < len(d):
            dt, i = _vlq(d, i)
            tick += dt
            st = d[i]
            if st & 0x80: i += 1
            else: st = rs                               # running status
            if st == 0xff:
                i += 1
                l, i = _vlq(d, i); i += l
            elif st in (0xf0, 0xf7):
                l, i = _vlq(d, i); i += l
            else:
                rs = st
                hi, ch = st & 0xf0, st & 0xf
                if hi in (0xc0, 0xd0):
                    if hi == 0xc0: ev.append(('prog', tick, ch, d[i]))
                    i += 1
                else:
                    a, v = d[i], d[i + 1]; i += 2
                    if hi == 0x90 and v > 0: ev.append(('on', tick, ch, a))
                    elif hi == 0x80 or (hi == 0x90 and v == 0): ev.append(('off', tick, ch, a))
        tracks.append(ev)
    return div, tracks


def _quantize(q, divisors=(4, 3)):
    "music21 Stream.quantize: snap to the nearest multiple of 1/d over d in divisors (first best wins)."
    best, best_err = None, None
    for d in divisors:
        cand = Fraction(round(q * d), d)
        err = abs(cand - q)
        if best is None or err < best_err:
            best, best_err = cand, err
    return best


# GM program -> the class the reference's instrument test (deep_music_genre.py:251-291) would land on, given
# music21's instrumentFromMidiProgram class tree.  None = "instrument rejected" (the `break` at :281).
def gm_program_category(prog):
    if prog <= 8 or prog == 55 or 80 <= prog <= 103 or prog >= 117: return 'Piano'     # KeyboardInstrument (Sampler incl.)
    if prog == 15: return 'StringInstrument'                                           # Dulcimer
    if 9 <= prog <= 23: return None                                                    # pitched percussion, organs
    if 24 <= prog <= 31: return 'Guitar'
    if prog == 32: return 'StringInstrument'                                           # AcousticBass(StringInstrument)
    if 33 <= prog <= 39: return 'Bass'                                                 # ElectricBass/FretlessBass(Guitar)
    if prog == 47: return None                                                         # Timpani
    if 40 <= prog <= 51: return 'StringInstrument'
    if 52 <= prog <= 54: return None                                                   # voices
    if 56 <= prog <= 63: return 'BrassInstrument'
    if 64 <= prog <= 79: return 'WoodwindInstrument'
    if 104 <= prog <= 107 or prog == 110: return 'StringInstrument'
    if prog in (109, 111): return 'WoodwindInstrument'
    return None


def smf_parts(path):
    """What ``stream2chordarr`` sees: (parts, highest_time) with parts = list over note-bearing tracks of a
    time-ordered element list: ('ins', offset_q, program) or ('note', offset_q, pitch, quarterLength)."""
    tpq, tracks = read_smf(path)
    parts, highest = [], Fraction(0)
    for ev in tracks:
        open_notes, notes, progs = {}, [], []
        for e in ev:
            if e[0] == 'prog':
                progs.append((Fraction(e[1], tpq), e[3]))
            elif e[0] == 'on':
                open_notes.setdefault((e[2], e[3]), []).append(e[1])
            else:
                q = open_notes.get((e[2], e[3]))
                if q:
                    t0 = q.pop(0)                                              # FIFO pairing
                    notes.append((t0, e[3], e[1] - t0))
        if not notes:
            continue
        elems = []
        for off, prog in progs:
            elems.append((_quantize(off), 0, ('ins', prog)))
        for t0, pitch, dur in notes:
            off, ql = _quantize(Fraction(t0, tpq)), _quantize(Fraction(dur, tpq))
            elems.append((off, 1, ('note', pitch, ql)))
            highest = max(highest, off + ql)
        elems.sort(key=lambda x: (x[0], x[1]))                                 # stable: file order within ties
        parts.append([(k[0], off) + tuple(k[1:]) for off, _, k in elems])
    return parts, highest


def stream2chordarr(parts, highest_time, note_size=NOTE_SIZE, sample_freq=SAMPLE_FREQ, max_note_dur=MAX_NOTE_DUR):
    "deep_music_genre.py:220-322 over the SMF-derived parts"
    maxTimeStep = round(highest_time * sample_freq) + 1
    score_arr = np.zeros((maxTimeStep, len(parts), note_size))
    ins = dict()
    for idx, part in enumerate(parts):
        notes, iterate = [], False
        for elem in part:
            if elem[0] == 'ins':
                cat = gm_program_category(elem[2])
                if cat is not None:
                    ins[idx] = cat
                    iterate = True
                else:
                    break
            else:
                _, off, pitch, ql = elem
                notes.append((pitch, int(round(off * sample_freq)), int(round(ql * sample_freq))))
        notes_sorted = sorted(notes, key=lambda x: (x[1], x[2]))
        if iterate:
            for pitch, offset, duration in notes_sorted:
                if max_note_dur is not None and duration > max_note_dur: duration = max_note_dur
                score_arr[offset, idx, pitch] = duration
                score_arr[offset + 1:offset + duration, idx, pitch] = VALTCONT
    return score_arr, ins


def timestep2npenc(timestep, note_range=NOTE_RANGE):
    "deep_music_genre.py:367-387 (enc_type='full')"
    notes = []
    for i, n in zip(*timestep.nonzero()):
        d = timestep[i, n]
        if d < 0: continue
        if n < note_range[0] or n >= note_range[1]: continue
        notes.append([n, d, i])
    notes = sorted(notes, key=lambda x: x[0], reverse=True)
    return [[n, d, i] for n, d, i in notes]


def chordarr2npenc(chordarr, skip_last_rest=True):
    "deep_music_genre.py:324-345"
    result, wait_count = [], 0
    for idx, timestep in enumerate(chordarr):
        flat_time = timestep2npenc(timestep)
        if len(flat_time) == 0:
            wait_count += 1
        else:
            if wait_count > 0: result.append([VALTSEP, wait_count, NI_NPENC])
            result.extend(flat_time)
            wait_count = 1
    if wait_count > 0 and not skip_last_rest: result.append([VALTSEP, wait_count, NI_NPENC])
    return np.array(result, dtype=int)


def sort_instruments(npenc):
    "deep_music_genre.py:1443-1487 (including the re-use of the last zip pair's separator for the tail group)"
    sep_idxs = (npenc[:, 0] == -1).nonzero()[0]
    updated = []
    first_sep = sep_idxs[0]
    if first_sep != 0:
        updated.extend(sorted(npenc[0:first_sep], key=lambda x: x[2]))
    e = None
    for e in zip(sep_idxs[:-1], sep_idxs[1:]):
        sub = sorted(npenc[e[0] + 1:e[1]], key=lambda x: x[2])
        sep = npenc[e[0]]
        updated.extend([sep] + sub)
    last_sep = sep_idxs[-1]
    if len(npenc) > last_sep + 1:
        sub = sorted(npenc[last_sep + 1:], key=lambda x: x[2])
        sep = npenc[e[0]]
        final_subset = [sep] + sub
    else:
        final_subset = [sep]
    updated.extend(final_subset)
    updated = np.array(updated)
    assert list(sep_idxs) == list((updated[:, 0] == -1).nonzero()[0])
    return updated


def npins2vocabins(x, ins):
    "deep_music_genre.py:1301-1312"
    if x in ins.keys():
        return ACCEP_INS[ins[x]] if ins[x] in ACCEP_INS.keys() else ACCEP_INS['Piano']
    elif x == NI_NPENC:
        return x
    raise Exception


def seq_prefix(vocab, genre=None):
    "deep_music_genre.py:1361-1376 (Sentence / Genre)"
    token = BOS
    if genre is not None:
        g = genre.lower()
        for key, tok in (('electronic', ELECTRONIC), ('folk', FOLK), ('funk', FUNK), ('jazz', JAZZ), ('pop', POP),
                         ('rock', ROCK)):
            if key in g:
                token = tok
                break
    return np.array([vocab.stoi[token], vocab.pad_idx])


def npenc2idxenc(t, vocab, ins=None, genre=None, add_eos=True):
    "deep_music_genre.py:1315-1359 (3-column branch)"
    t = t.copy()
    t[:, 0] = t[:, 0] + vocab.note_range[0]
    t[:, 1] = t[:, 1] + vocab.dur_range[0]
    if ins is not None:
        t[:, 2] = np.array([npins2vocabins(x, ins) for x in t[:, 2]])
    t[:, 2] = t[:, 2] + vocab.ins_range[0]
    prefix = seq_prefix(vocab, genre)
    suffix = np.array([vocab.stoi[EOS]]) if add_eos else np.empty(0, dtype=int)
    return np.concatenate([prefix, t.reshape(-1), suffix])


def position_enc(idxenc, vocab):
    "deep_music_genre.py:1489-1527"
    sep_idxs = (idxenc == vocab.sep_idx).nonzero()[0]
    sep_idxs = sep_idxs[sep_idxs + 2 < idxenc.shape[0]]
    dur_vals = idxenc[sep_idxs + 1]
    dur_vals[dur_vals == vocab.mask_idx] = vocab.dur_range[0]
    dur_vals -= vocab.dur_range[0]
    posenc = np.zeros_like(idxenc)
    try:
        if len(idxenc) > sep_idxs[-1] + 3:
            posenc[sep_idxs + 3] = dur_vals
        else:
            sep_idxs = sep_idxs[:-1]
            dur_vals = dur_vals[:-1]
            posenc[sep_idxs + 3] = dur_vals
    except Exception:
        pass                                                    # reference prints and carries on (:1516-1525)
    return posenc.cumsum()


def find_beat(pos, beat, sample_freq=SAMPLE_FREQ, side='left'):
    return np.searchsorted(pos, beat * sample_freq, side=side)


def beat2index(idxenc, pos, vocab, beat, include_last_sep=False):
    "deep_music_genre.py:1529-1537"
    cutoff = find_beat(pos, beat)
    if cutoff < 2: return 2
    if len(idxenc) < 2 or include_last_sep: return cutoff
    if idxenc[cutoff - 2] == vocab.sep_idx: return cutoff - 2
    return cutoff


def trim_to_beat(idxenc, pos, vocab, to_beat=None, include_last_sep=True):
    "deep_music_genre.py:1546-1549"
    if to_beat is None: return idxenc
    return idxenc[:beat2index(idxenc, pos, vocab, to_beat, include_last_sep=include_last_sep)]


def midi_to_idxenc(path, vocab, genre=None):
    "MusicItem.from_file (deep_music_genre.py:1167-1195): file -> chordarr -> npenc -> sorted -> idxenc"
    parts, highest = smf_parts(path)
    chordarr, ins = stream2chordarr(parts, highest)
    npenc = chordarr2npenc(chordarr)
    npenc = sort_instruments(npenc)
    return npenc2idxenc(npenc, vocab, ins=ins, genre=genre)


def seed_from_midi(path, vocab, cutoff_beat=None, genre_token=None, strip_eos=True):
    """The seed construction of ``predictNwGenreModel`` (app_utils.py:112-126) / notebook cells 76-78:
    ``MusicItem.from_file(..).trim_to_beat(cutoff_beat)`` (MusicItem.trim_to_beat passes
    include_last_sep=False, deep_music_genre.py:1244-1245), ``data[0] = stoi[genre_token]``, drop a trailing xxeos."""
    idx = midi_to_idxenc(path, vocab)
    if cutoff_beat is not None:
        idx = trim_to_beat(idx, position_enc(idx, vocab), vocab, cutoff_beat, include_last_sep=False)
    idx = idx.copy()
    if genre_token is not None:
        idx[0] = vocab.stoi[genre_token]
    if strip_eos and len(idx) and idx[-1] == vocab.stoi[EOS]:
        idx = idx[:-1]
    return idx
