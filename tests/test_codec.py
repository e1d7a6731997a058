"""Host codec of the product (deepmusicgeneration_b200.codec) against the reference golden and the oracle restatement."""
import os

import numpy as np
import pytest

from deepmusicgeneration_b200 import codec as pc
from oracle import codec as oc

MIDIS = ['Undertale_-_Megalovania.mid', 'fur_elise.mid', 'uploadedMidi.mid', 'Never_Gonna_Let_You_Go.mid']


def test_vocab_identical_to_reference_layout():
    v, o = pc.MusicVocab.create(), oc.MusicVocab.create()
    assert v.itos == o.itos and len(v) == 324
    for name in ('mask_idx', 'pad_idx', 'bos_idx', 'sep_idx', 'ni_idx', 'npenc_range', 'note_range', 'dur_range', 'ins_range'):
        assert getattr(v, name) == getattr(o, name)
    for idx in range(324):
        assert bool(v.is_duration(idx)) == bool(o.is_duration(idx)) and bool(v.is_note(idx)) == bool(o.is_note(idx))
        assert bool(v.is_ins(idx)) == bool(o.is_ins(idx)) and bool(v.is_duration_or_pad(idx)) == bool(o.is_duration_or_pad(idx))


def test_megalovania_golden_bit_exact(golden_dir):
    v = pc.MusicVocab.create()
    gold = open(os.path.join(golden_dir, 'megalovania_seed64.txt')).read().split()
    item = pc.MusicItem.from_file(os.path.join(golden_dir, MIDIS[0]), v).trim_to_beat(64)
    item.data[0] = v.stoi['xxelec']
    assert item.to_text().split(' ') == gold


@pytest.mark.parametrize('name', MIDIS)
def test_product_codec_equals_oracle(golden_dir, name):
    v, o = pc.MusicVocab.create(), oc.MusicVocab.create()
    path = os.path.join(golden_dir, name)
    a = pc.MusicItem.from_file(path, v)
    b = oc.midi_to_idxenc(path, o)
    assert np.array_equal(a.data, b)
    assert np.array_equal(a.position, oc.position_enc(b, o))
    for beat in (0, 1, 8, 32, 64, 10 ** 6):
        ta = a.trim_to_beat(beat).data
        tb = oc.trim_to_beat(b, oc.position_enc(b, o), o, beat, include_last_sep=False)
        assert np.array_equal(ta, tb), beat


def test_idxenc_roundtrip_and_midi_writer(golden_dir, tmp_path):
    v = pc.MusicVocab.create()
    item = pc.MusicItem.from_file(os.path.join(golden_dir, 'uploadedMidi.mid'), v)
    npenc = item.to_npenc()
    assert npenc.shape[1] == 3 and (npenc[:, 1] >= 0).all()
    out = tmp_path / 'out.mid'
    item.to_stream(bpm=120).write('midi', fp=str(out))
    again = pc.MusicItem.from_file(str(out), v)
    assert np.array_equal(item.data, again.data)


def test_edge_cases():
    v = pc.MusicVocab.create()
    empty = pc.MusicItem.empty(v)
    assert list(empty.data) == [v.bos_idx, v.pad_idx] and list(empty.position) == [0, 0]
    assert list(pc.trim_to_beat(empty.data, empty.position, v, 4, include_last_sep=False)) == [v.bos_idx, v.pad_idx]
    masked = pc.MusicItem(np.array([0, 1, 72, 142, 301, 11, 142, 10, 74, 144, 301]), v).mask_pitch()
    assert list(masked.data) == [0, 1, 3, 142, 301, 11, 142, 10, 3, 144, 301]
    assert list(masked.position) == [0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2]
    t = pc.MusicItem(np.array([0, 1, 72, 142, 301]), v).transpose(2)
    assert list(t.data) == [0, 1, 74, 142, 301]


@pytest.mark.parametrize('name', ['ref_genre_output.mid', 'ref_remix_Notes_output.mid'])
def test_midi_writer_reproduces_the_reference_outputs_byte_for_byte(golden_dir, tmp_path, name):
    """outputs/genre_output.mid and outputs/remix_Notes_output.mid are files the reference itself wrote with
    ``full.to_stream(bpm).write('midi', fp)`` (app.py:191, 275; music21 behind it).  Decoding them and writing them again with the
    product's SMF writer must give the same bytes: tempo / key / time-signature track, track name + program + pitch-bend prelude,
    1024 ticks per quarter, velocity 90, note-off ordering, chord grouping by duration (deep_music_genre.py:513-541), end-of-track gaps."""
    v = pc.MusicVocab.create()
    src = os.path.join(golden_dir, name)
    item = pc.MusicItem.from_file(src, v)
    out = tmp_path / 'again.mid'
    item.to_stream(bpm=120).write('midi', fp=str(out))
    assert open(src, 'rb').read() == open(out, 'rb').read()
