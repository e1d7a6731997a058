"""The Streamlit app's glue functions (reference ``app_utils.py:13-215``) on the B200 engine: same names, same
arguments, same return values, so ``app.py:179-191, 264-275`` keeps working when it imports this module instead."""
import numpy as np

from .codec import MusicDataBunch, MusicItem, MusicVocab
from .learner import multitask_model_learner, music_model_learner
from .model import tfmerXL_lm_config

_GENRE_PREFIX = (('pop', 'xxpop'), ('folk', 'xxfolk'), ('jazz', 'xxjazz'), ('rock', 'xxrock'), ('funk', 'xxfunk'),
                 ('elec', 'xxelec'))


def default_config():
    config = tfmerXL_lm_config()
    config.update(act='gelu', mem_len=512, d_model=512, d_inner=2048, n_layers=6, n_heads=8, d_head=64)
    return config


def music_config():
    config = default_config()
    config['ctx_len'] = 512
    return config


def btp_phase1_config():
    config = default_config()
    config.update(ctx_len=512, d_inner=3072, n_heads=12, d_head=64, n_layers=8, transpose_range=(0, 12), mask_steps=4,
                  encode_position=False)
    return config


def multitask_config():
    config = music_config()
    config.update(encode_position=True, bias=True, enc_layers=10, dec_layers=10)
    del config['n_layers']
    return config


def baseline_config():
    "BASELINE.json 'musicautobot default': d_model 512, 16 layers, 8 heads, mem_len 512."
    config = default_config()
    config.update(ctx_len=512, n_layers=16, encode_position=False, mask_steps=1)
    return config


def createGenreContinuationModel(encode_position=False, ckpt_path='./checkpoints/lakh_genre_model.pth', **engine_kwargs):
    config = btp_phase1_config()
    return music_model_learner(MusicDataBunch.empty(''), config=config.copy(), encode_position=encode_position,
                               pretrained_path=ckpt_path, **engine_kwargs)


def createRemixModel(encode_position=True, ckpt_path='./checkpoints/mask_music_model.pth', **engine_kwargs):
    config = multitask_config()
    return multitask_model_learner(MusicDataBunch.empty(''), config=config.copy(), pretrained_path=ckpt_path,
                                   **engine_kwargs)


def _genre_prefix(genre):
    genre = genre.lower().strip()
    for key, tok in _GENRE_PREFIX:
        if key in genre:
            return tok
    return None


def _seed_item(mid_file, genre, cutoff_beat):
    vocab = MusicVocab.create()
    item = MusicItem.from_file(mid_file, vocab)
    seed_item = item.trim_to_beat(cutoff_beat)
    prefix = _genre_prefix(genre)
    if prefix is not None:
        seed_item.data[0] = vocab.stoi[prefix]
    else:
        seed_item.data = seed_item.data[1:]
    if seed_item.to_text().split(' ')[-1] == 'xxeos':
        seed_item.data = seed_item.data[:len(seed_item.data) - 1]
    return vocab, seed_item


def predictNwGenreModel(genre_model_learner, mid_file, genre=' POP ', temperature_notes=1.8, temperature_duration=1.8,
                        temperature_ins=1.0, top_p=0.3, max_len=512, cutoff_beat=32, mem_len=512, allowed_ins=[],
                        output_bpm=120):
    "app_utils.py:90-144 (note: like the reference, top_p is not forwarded - predict runs with top_k=30, top_p=0.65)"
    genre_model_learner.model.mem_len = mem_len          # as in the reference, this does not reach model[0].mem_len
    _, seed_item = _seed_item(mid_file, genre, cutoff_beat)
    if allowed_ins == []:
        allowed_ins = None
    else:
        rename = {'Flute': 'WoodwindInstrument', 'Brass': 'BrassInstrument', 'Violin': 'StringInstrument'}
        for idx, ins in enumerate(allowed_ins):
            allowed_ins[idx] = rename.get(ins, ins)
    pred, full = genre_model_learner.predict(seed_item, n_words=max_len,
                                             temperatures=(temperature_notes, temperature_duration, temperature_ins),
                                             min_bars=12, top_k=30, top_p=0.65, allowed_ins=allowed_ins)
    return full


def predictMaskModel(mask_model_learner, mid_file, genre=' POP ', temperature_notes=1.0, temperature_duration=1.0, top_p=0.3,
                     cutoff_beat=32, output_bpm=120, pred_type='notes', mask_proportion=0.6):
    "app_utils.py:159-215"
    vocab, seed_item = _seed_item(mid_file, genre, cutoff_beat)
    first = 'n' if pred_type == 'notes' else 'd'
    indices = [i for i, x in enumerate(vocab.textify(seed_item.data).split(' ')) if x[0] == first]
    selected = np.random.choice(indices, int(len(indices) * mask_proportion), replace=False)
    seed_item.data[selected] = vocab.mask_idx
    if pred_type == 'notes':
        return mask_model_learner.predict_mask(seed_item, temperatures=(temperature_notes, temperature_duration))
    return mask_model_learner.predict_mask(seed_item, temperatures=(0.8, 0.8), top_k=40, top_p=0.6)
