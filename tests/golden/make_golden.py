"""Regenerates the fixtures under tests/golden/ from the read-only reference checkout.

Run in the build container only (needs /root/reference): ``python tests/golden/make_golden.py``.
* ``megalovania_seed64.txt``: the executed-notebook output of ``seed_item.to_text()``
  (notebooks/Transformer_Genre_Evaluation.ipynb cell 79, 623 tokens) = MusicItem.from_file(
  'Undertale_-_Megalovania.mid').trim_to_beat(64) with data[0] = stoi['xxelec'].
* ``notebook_pins.json``: the scalar known answers (vocab size, parameter count, xxni index).
* the seed MIDI files themselves (inputs, not source code), so that the GPU box - which has no
  /root/reference - can run the codec tests.
"""
import json, os, re, shutil

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))

nb = json.load(open(os.path.join(REF, 'notebooks/Transformer_Genre_Evaluation.ipynb')))
cells = nb['cells']


def out_text(cell):
    for o in cell.get('outputs', []):
        t = o.get('text') or o.get('data', {}).get('text/plain')
        if t:
            return ''.join(t)
    return ''


seed = None
for c in cells:
    if ''.join(c['source']).strip() == 'seed_item.to_text()':
        seed = out_text(c).strip().strip("'")
assert seed is not None and len(seed.split(' ')) == 623, len(seed.split(' '))
open(os.path.join(HERE, 'megalovania_seed64.txt'), 'w').write(seed + '\n')

pins = {}
for c in cells:
    src = ''.join(c['source'])
    if src.strip() == 'calc_net_weight_count(learner.model)':
        pins['btp_phase1_param_count'] = int(out_text(c).strip())
    if 'print(len(data_vocab))' in src:
        pins['vocab_size'] = int(out_text(c).strip())
    if 'learner.predict(seed_item' in src:
        for o in c['outputs']:
            t = ''.join(o.get('text', ''))
            m = re.search(r'Init prev_idx =\s+(\d+)', t)
            if m: pins['init_prev_idx'] = int(m.group(1))
json.dump(pins, open(os.path.join(HERE, 'notebook_pins.json'), 'w'), indent=1)
print(pins)

for f in ['Undertale_-_Megalovania.mid', 'fur_elise.mid', 'tempDir/uploadedMidi.mid', 'Never_Gonna_Let_You_Go.mid']:
    shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, os.path.basename(f)))
    os.chmod(os.path.join(HERE, os.path.basename(f)), 0o644)
