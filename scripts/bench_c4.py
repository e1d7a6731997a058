"C4 (BASELINE.json configs[3]): masked-BERT remix encoder forward, seq 1024, batch 512, bf16 - forward tokens/s (CUDA events)."
import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200.model import get_multitask_model
from deepmusicgeneration_b200.app_utils import multitask_config
from deepmusicgeneration_b200 import _lib
B = int(os.environ.get('C4_BATCH', 512)); T = 1024; V = 324
cfg = multitask_config()
chunk = 32                                             # streams per activation chunk (32 x 1024 rows)
pm = get_multitask_model(V, cfg, pad_idx=1, dtype='bf16', max_batch=B, max_seq=T, max_rows=chunk * T, seed=0)
g = torch.Generator().manual_seed(1234)
x = torch.randint(0, V, (B, T), generator=g).cuda()
pos = torch.cumsum(torch.randint(0, 9, (B, T), generator=g), 1).clamp_max(32 * 1024 - 1).cuda()
e = pm._e
def fwd():
    return e.forward(x, pos, _lib.LOGITS_NONE)
fwd(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get('C4_REPS', 3))
e0.record()
for _ in range(reps): fwd()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
L = cfg['enc_layers']
flops_tok = L * (3 * 2 * 512 * 512 + 2 * 512 * 512 * 0 + 3 * 2 * 512 * T)      # q,k,v GEMMs + AC, BD, PV (dense count)
print(json.dumps({'workload': f'C4: remix encoder {L} layers d 512, 8 heads, seq {T}, batch {B}, bf16 forward (no logits)',
                  'flash': not bool(os.environ.get('DMG_NO_FLASH')), 'ms_per_forward': ms, 'tokens_per_s': B * T / (ms / 1e3),
                  'tflops_dense_count': flops_tok * B * T / (ms / 1e3) / 1e12}))
