"""Oracle: the reference's token-by-token generation loops.  TEST INFRASTRUCTURE.

Restates ``top_k_top_p`` (``deep_music_genre.py:1679-1706``), ``filter_invalid_indexes``
(``:1984-2018``; remix variant ``deep_music_remix.py:2394-2437``), ``MusicLearner.predict``
(``deep_music_genre.py:1853-1972``) and ``MultitaskLearner.predict_mask``
(``deep_music_remix.py:2563-2613``) over the oracle models, on the CPU, batch size 1, one host
round trip per token - exactly the reference's schedule.

Determinism: ``torch.multinomial`` draws from torch's global generator.  The "greedy" stream that
``north_star`` wants bit-exact is this loop with ``top_k=1, top_p=0.0`` (one finite logit survives
``top_k_top_p``, so ``multinomial`` has a single choice) - SURVEY.md 8(c).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .codec import (SPECIAL_TOKS, BOS, PAD, EOS, SEP, IN, ELECTRONIC, FOLK, FUNK, JAZZ, POP, ROCK, ACCEP_INS,
                    SAMPLE_FREQ)


def top_k_top_p(logits, top_k=0, top_p=0.0, filter_value=-float('Inf')):
    "deep_music_genre.py:1679-1706"
    logits = logits.clone()
    assert logits.dim() == 1
    top_k = min(top_k, logits.size(-1))
    if top_k > 0:
        indices_to_remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
        logits[indices_to_remove] = filter_value
    if top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        sorted_indices_to_remove = cumulative_probs > top_p
        sorted_indices_to_remove[..., 1:] = sorted_indices_to_remove[..., :-1].clone()
        sorted_indices_to_remove[..., 0] = 0
        indices_to_remove = sorted_indices[sorted_indices_to_remove]
        logits[indices_to_remove] = filter_value
    return logits


def filter_invalid_indexes(res, prev_idx, vocab, filter_value=-float('Inf'), last_xxsep=False, allowed_ins=None):
    "deep_music_genre.py:1984-2018 (note: mt*/dummy* tokens are never filtered by the reference)"
    if allowed_ins is not None:
        res[list(set(range(vocab.ins_range[0], vocab.ins_range[1])) - set([vocab.stoi[x] for x in allowed_ins]))] = filter_value
    if last_xxsep is True:
        res[list(range(*vocab.ins_range))] = filter_value
    else:
        res[[vocab.stoi[IN]]] = filter_value
    if vocab.is_duration(prev_idx):
        res[list(range(*vocab.dur_range))] = filter_value
        res[list(range(*vocab.note_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]) - {vocab.stoi[IN]})] = filter_value
    elif vocab.is_ins(prev_idx) or prev_idx == vocab.stoi[PAD]:
        res[list(range(*vocab.ins_range))] = filter_value
        res[list(range(*vocab.dur_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]) - {vocab.stoi[SEP]})] = filter_value
    else:
        res[list(range(*vocab.note_range))] = filter_value
        res[list(range(*vocab.ins_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]))] = filter_value
    return res


def filter_invalid_indexes_remix(res, prev_idx, vocab, filter_value=-float('Inf'), last_xxsep=False, allowed_ins=None):
    "deep_music_remix.py:2394-2437"
    if allowed_ins is not None:
        res[list(set(range(vocab.ins_range[0], vocab.ins_range[1]))
                 - set([vocab.stoi['i' + str(ACCEP_INS[x])] for x in allowed_ins]))] = filter_value
    if last_xxsep is True:
        res[list(range(*vocab.ins_range))] = filter_value
    else:
        res[[vocab.stoi[IN]]] = filter_value
    if vocab.pad_idx == prev_idx:
        res[list(range(*vocab.dur_range))] = filter_value
        res[list(range(*vocab.ins_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]) - {vocab.stoi[SEP]})] = filter_value
    elif vocab.is_duration(prev_idx):
        res[list(range(*vocab.dur_range))] = filter_value
        res[list(range(*vocab.note_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]) - {vocab.stoi[IN]})] = filter_value
    elif vocab.is_ins(prev_idx):
        res[list(range(*vocab.ins_range))] = filter_value
        res[list(range(*vocab.dur_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]) - {vocab.stoi[SEP]})] = filter_value
    else:
        res[list(range(*vocab.note_range))] = filter_value
        res[list(range(*vocab.ins_range))] = filter_value
        res[list(set([vocab.stoi[x] for x in SPECIAL_TOKS]))] = filter_value
    return res


def predict(model, vocab, item_data, item_pos, n_words=128, temperatures=(1.0, 1.0, 1.0), min_bars=4, top_k=30,
            top_p=0.6, allowed_ins=None, verbose=False, on_step=None):
    """``MusicLearner.predict`` (deep_music_genre.py:1853-1972) over an oracle ``SequentialRNN``.

    ``item_data``/``item_pos`` are the seed's idxenc and position arrays (``MusicItem.data``/``.position``).
    Returns the list of generated token ids (``new_idx``).  ``on_step(i, logits)`` is a test hook that sees
    the raw last-position logits before temperature/filtering.
    """
    model.reset()
    new_idx = []
    x = torch.as_tensor(np.asarray(item_data)).long()
    pos = torch.as_tensor(np.asarray(item_pos)).long()
    last_pos = pos[-1] if len(pos) else 0
    start_pos = last_pos
    repeat_count = 0
    encode_position = getattr(model[0], 'encode_position', False)
    last_xxsep = False
    if allowed_ins is not None:
        for idx, ins in enumerate(allowed_ins):
            allowed_ins[idx] = 'i' + str(ACCEP_INS[ins])

    for i in range(n_words):
        with torch.no_grad():
            if encode_position:
                batch = {'x': x[None], 'pos': pos[None]}
                logits = model(batch)[0][-1][-1]
            else:
                logits = model(x[None])[0][-1][-1]
        if on_step is not None:
            on_step(i, logits.clone())

        if len(new_idx):
            prev_idx = new_idx[-1]
        else:
            prev_idx = int(item_data[-1])
            if verbose: print('Init prev_idx = ', prev_idx)

        if prev_idx == vocab.sep_idx:
            last_xxsep = True
        elif vocab.is_ins(prev_idx):
            if prev_idx == vocab.ni_idx:
                last_xxsep = False

        temperature = None
        if vocab.is_duration(prev_idx):
            temperature = temperatures[2]
        elif vocab.is_note(prev_idx):
            temperature = temperatures[1]
        elif vocab.is_ins(prev_idx) or prev_idx == vocab.stoi[PAD]:
            temperature = temperatures[0]
        if temperature is None:
            raise AssertionError(f'Assertion error: prev_idx = {vocab.itos[prev_idx]}')

        repeat_penalty = max(0, np.log((repeat_count + 1) / 4) / 5) * temperature
        temperature += repeat_penalty
        if temperature != 1.:
            logits = logits / temperature

        filter_value = -float('Inf')
        if ((last_pos - start_pos) // 16) <= min_bars:
            logits[vocab.bos_idx] = filter_value

        logits = filter_invalid_indexes(logits, prev_idx, vocab, filter_value=filter_value, last_xxsep=last_xxsep,
                                        allowed_ins=allowed_ins)
        logits = top_k_top_p(logits, top_k=top_k, top_p=top_p, filter_value=filter_value)

        probs = F.softmax(logits, dim=-1)
        idx = torch.multinomial(probs, 1).item()

        num_choices = len(probs.nonzero().view(-1))
        if num_choices <= 2:
            repeat_count += 1
        else:
            repeat_count = repeat_count // 2

        if prev_idx == vocab.sep_idx:
            duration = idx - vocab.dur_range[0]
            last_pos = last_pos + duration
            abs_bar = last_pos // 16
            if (i / n_words > 0.80) and (abs_bar % 4 == 0):
                break

        if idx == vocab.bos_idx:
            if verbose: print('Predicted BOS token. Returning prediction...')
            break

        new_idx.append(idx)
        x = x.new_tensor([idx])
        pos = pos.new_tensor([int(last_pos)])
    return new_idx


def beam_search(model, xb, n_words, top_k=10, beam_sz=10, temperature=1., return_beams=False):
    """``MusicLearner.beam_search`` (deep_music_genre.py:1823-1851) over an oracle ``SequentialRNN``.  ``argsort`` is made stable
    (the reference's is not, and its top_k identical copies of the seed produce exact ties); ``return_beams`` returns the surviving
    (token history, score) pairs instead of drawing one of them."""
    model.reset()
    model.eval()
    xb_length = xb.shape[-1]
    if xb.shape[0] > 1: xb = xb[0][None]
    xb = xb.repeat(top_k, 1)
    nodes = xb.clone()
    scores = xb.new_zeros(1).float()
    with torch.no_grad():
        for k in range(n_words):
            out = F.log_softmax(model(xb)[0][:, -1], dim=-1)
            values, indices = out.topk(top_k, dim=-1)
            scores = (-values + scores[:, None]).view(-1)
            indices_idx = torch.arange(0, nodes.size(0))[:, None].expand(nodes.size(0), top_k).contiguous().view(-1)
            sort_idx = scores.argsort(stable=True)[:beam_sz]
            scores = scores[sort_idx]
            nodes = torch.cat([nodes[:, None].expand(nodes.size(0), top_k, nodes.size(1)),
                               indices[:, :, None].expand(nodes.size(0), top_k, 1)], dim=2)
            nodes = nodes.view(-1, nodes.size(2))[sort_idx]
            model[0].select_hidden(indices_idx[sort_idx])
            xb = nodes[:, -1][:, None]
    if temperature != 1.: scores.div_(temperature)
    if return_beams:
        return nodes[:, xb_length:].numpy(), scores
    node_idx = torch.multinomial(torch.exp(-scores), 1).item()
    return [i.item() for i in nodes[node_idx][xb_length:]]


def predict_mask(model, vocab, item_data, item_pos, temperatures=(1.0, 1.0), top_k=20, top_p=0.8):
    "``MultitaskLearner.predict_mask`` (deep_music_remix.py:2563-2613) over an oracle ``MultiTransformer`` (eval mode)."
    x = torch.as_tensor(np.asarray(item_data)).long().clone()
    pos = torch.as_tensor(np.asarray(item_pos)).long()
    model.eval()
    model.reset()
    mask_idxs = (x == vocab.mask_idx).nonzero().view(-1)
    repeat_count = 0
    for midx in mask_idxs:
        prev_idx = x[midx - 1]
        with torch.no_grad():
            logits = model({'msk': {'x': x[None], 'pos': pos[None]}})['msk'][0][midx]
        temperature = temperatures[0] if vocab.is_duration_or_pad(prev_idx) else temperatures[1]
        repeat_penalty = max(0, np.log((repeat_count + 1) / 4) / 5) * temperature
        temperature += repeat_penalty
        if temperature != 1.:
            logits = logits / temperature
        filter_value = -float('Inf')
        special_idxs = [vocab.bos_idx, vocab.sep_idx, vocab.stoi[IN], vocab.stoi[EOS], vocab.stoi[ELECTRONIC],
                        vocab.stoi[FOLK], vocab.stoi[FUNK], vocab.stoi[JAZZ], vocab.stoi[POP], vocab.stoi[ROCK]]
        logits[special_idxs] = filter_value
        logits = filter_invalid_indexes_remix(logits, prev_idx, vocab, filter_value=filter_value)
        logits = top_k_top_p(logits, top_k=top_k, top_p=top_p, filter_value=filter_value)
        probs = F.softmax(logits, dim=-1)
        idx = torch.multinomial(probs, 1).item()
        num_choices = len(probs.nonzero().view(-1))
        if num_choices <= 2:
            repeat_count += 1
        else:
            repeat_count = repeat_count // 2
        x[midx] = idx
    return x.cpu().numpy()
