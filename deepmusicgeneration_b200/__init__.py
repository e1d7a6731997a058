"""B200-native (sm_100a) Transformer-XL / masked-BERT hot path of DeepMusicGeneration.

Host code is Python/PyTorch (device buffers, streams) over the C ABI of ``include/dmg_b200.h``
(``libdmg_b200.so``: hand-written CUDA - tcgen05/TMEM GEMMs fed by TMA, fused relative-position attention over a
K/V ring in HBM, warp-shuffle LayerNorm, fused sampling).  No CPU fallback: the CUDA library is required.
"""
from .codec import (ACCEP_INS, MusicDataBunch, MusicItem, MusicVocab, SEQType, idxenc2npenc, midi2npenc, npenc2idxenc,
                    position_enc, sort_instruments, trim_to_beat)

__all__ = ['ACCEP_INS', 'MusicDataBunch', 'MusicItem', 'MusicVocab', 'SEQType', 'idxenc2npenc', 'midi2npenc',
           'npenc2idxenc', 'position_enc', 'sort_instruments', 'trim_to_beat', 'music_model_learner',
           'multitask_model_learner', 'get_language_model', 'get_multitask_model']


def __getattr__(name):          # the engine-backed objects load libdmg_b200.so on first use
    if name in ('music_model_learner', 'multitask_model_learner', 'MusicLearner', 'MultitaskLearner', 'predict_from_midi'):
        from . import learner
        return getattr(learner, name)
    if name in ('get_language_model', 'get_multitask_model', 'init_state_dict'):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
