"""CPU oracle for the DeepMusicGeneration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepmusicgeneration_b200/`` may import,
call or link this package: it is the checker (``tests/``, ``__graft_entry__.smoke``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``), never the
thing measured or shipped.

What it restates (all citations relative to ``/root/reference``):

* ``oracle/txl.py``      - fastai==1.0.61 ``TransformerXL`` stack (un-vendored third-party
  dependency of the reference, pinned by ``notebooks/Transformer_Genre_Evaluation.ipynb:162``)
  plus the in-repo override ``MusicTransformerXL`` (``deep_music_genre.py:1577-1665``).
* ``oracle/bert.py``     - masked-BERT remix encoder + head (``deep_music_remix.py:1851-2104``).
* ``oracle/sampling.py`` - ``top_k_top_p`` / ``filter_invalid_indexes`` / ``MusicLearner.predict`` /
  ``predict_mask`` (``deep_music_genre.py:1679-1706, 1853-2018``, ``deep_music_remix.py:2394-2437, 2563-2613``).
* ``oracle/codec.py``    - vocab + MIDI->npenc->idxenc codec (``deep_music_genre.py:126-196, 220-387,
  812-890, 1301-1549``), with a plain Standard-MIDI-File reader standing in for music21.

Pinning status (see DESIGN.md "Oracle"):

* architecture: PINNED by the reference's known answer 41,107,268 parameters
  (``notebooks/Transformer_Genre_Evaluation.ipynb:3173``) - ``tests/test_oracle_pins.py``.
* vocab: PINNED (324, ``...ipynb:3223``; ``xxni`` = 10, ``...ipynb:3379``).
* MIDI->token: PINNED on the 623-token Megalovania golden (``...ipynb:3299``) committed under
  ``tests/golden``; fur_elise.mid (off-grid onsets, needs music21's quantiser) is PARITY UNPINNED.
* forward numerics (logits): PARITY UNPINNED at the fastai boundary - the reference holds no stored
  activations or checkpoints and fastai/music21 are not importable here (no network).  Logit parity is
  therefore CUDA-vs-this-restatement on identical random-init weights, as ``north_star`` defines it.
"""
