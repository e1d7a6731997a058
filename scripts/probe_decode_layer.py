"""Timeline of one fused one-token layer launch (decode_layer.cu) inside the C2 decode step: %globaltimer marks of CTA 0.
Run on the GPU box:  DMG_DECODE_TIMELINE=1 python scripts/probe_decode_layer.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('DMG_DECODE_TIMELINE', '1')
from deepmusicgeneration_b200 import _lib
from deepmusicgeneration_b200.app_utils import baseline_config
from deepmusicgeneration_b200.model import get_language_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = dict(baseline_config(), ctx_len=512)
m = get_language_model(324, cfg, dtype='bf16', device=0, max_batch=B, max_seq=512, max_rows=B * 64, keep_hidden=False, seed=0)
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 324, (B, 512), generator=g).cuda()
m.reset()
m._e.forward(x, None, _lib.LOGITS_LAST)
for _ in range(20):
    m._e.forward(torch.randint(0, 324, (B, 1), generator=g).cuda(), None, _lib.LOGITS_LAST)
torch.cuda.synchronize()
out = np.zeros(64, dtype=np.uint64)
_lib.check(m._e.lib.dmg_decode_timeline(m._e.h, out.ctypes.data_as(C.c_void_p)), 'timeline')
t0 = int(min(v for v in out[:48] if v > 0))
names = {0: ['start', 'A issued', 'B issued', 'C weights pre-issued', 'barrier 2 passed', 'C issued', 'D issued'],
         1: ['start', 'A mma issued', 'xa_ready0 seen', 'B mma issued', 'C kb0 issued', 'C mma issued', 'xa_ready1 seen', 'D mma issued', 'C kb7 issued'],
         2: ['start', 'pdl_wait done', 'tmem_full A', 'A epilogue done', 'barrier 1 passed', 'LN1 done', 'tmem_full B', 'B epilogue done', 'barrier 2 passed',
             'tmem_full C', 'C epilogue done', 'barrier 3 passed', 'LN2 done', 'tmem_full D', 'D epilogue done']}
for role, label in ((0, 'producer'), (1, 'mma'), (2, 'epilogue')):
    print(label)
    for i, n in enumerate(names[role]):
        v = int(out[role * 16 + i])
        if v: print(f'   {n:24s} {(v - t0) / 1e3:8.2f} us')

if out[48]:
    print('attention role of the same dual-role launch')
    for i, n in ((48, 'first CTA starts'), (49, 'rel-pos table built'), (50, 'first CTA ends'), (51, 'last CTA ends')):
        print(f'   {n:24s} {(int(out[i]) - t0) / 1e3:8.2f} us')
