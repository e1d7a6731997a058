"""Timeline of CTA 0 of one attn_bert_tc_kernel launch at the C4 geometry (32 sequences x 1024 tokens x 8 heads): %globaltimer marks of
softmax warps 0 / 4, the MMA issuer, the transform warps and the producer over the CTA's first four items.
Run on the GPU box:  python scripts/probe_bert_tc.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATH = os.environ.setdefault('DMG_BERT_TC_TIMELINE', '/tmp/bert_tc_timeline.bin')
from deepmusicgeneration_b200.model import get_multitask_model
from deepmusicgeneration_b200.app_utils import multitask_config

B, T = 32, 1024
cfg = dict(multitask_config())
m = get_multitask_model(324, cfg, pad_idx=1, dtype='bf16', max_batch=B, max_seq=T)
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 324, (B, T), generator=g).cuda()
pos = torch.cumsum(torch.randint(0, 9, (B, T), generator=g), 1).cuda()
for _ in range(3):
    m({'msk': {'x': x, 'pos': pos}})
torch.cuda.synchronize()
d = np.fromfile(PATH, dtype=np.uint64).astype(np.int64)
t0 = d[d > 0].min()
us = lambda v: f'{(v - t0) / 1e3:7.2f}' if v > 0 else '      -'
for item in range(4):
    print(f'item {item}')
    p = d[832 + 4 * item:832 + 4 * item + 3]
    tr = d[768 + 4 * item:768 + 4 * item + 2]
    print(f'  producer: q requested {us(p[0])}  first position-key blocks requested {us(p[1])}  V(0) requested {us(p[2])}')
    print(f'  transform: q arrived {us(tr[0])}  operand tiles ready {us(tr[1])}')
    for n in range(8):
        a = d[48 * item + 5 * n:48 * item + 5 * n + 5]
        b = d[256 + 48 * item + 5 * n:256 + 48 * item + 5 * n + 5]
        i = d[512 + 48 * item + 6 * n:512 + 48 * item + 6 * n + 5]
        w = d[1024 + 64 * item + 8 * n:1024 + 64 * item + 8 * n + 8]
        print(f'  tile {n}: S issued {us(i[0])} PV issued {us(i[4])} | P out of the 8 warps ' + ' '.join(us(v) for v in w))
        print(f'           half 0: top {us(a[0])} S seen {us(a[1])} S read {us(a[2])} folded {us(a[3])} P out {us(a[4])}'
              f' | half 1: top {us(b[0])} S seen {us(b[1])} S read {us(b[2])} folded {us(b[3])} P out {us(b[4])}')
    print(f'  last PV seen {us(d[48 * item + 40])} / {us(d[256 + 48 * item + 40])}  item done {us(d[48 * item + 41])} / {us(d[256 + 48 * item + 41])}')
