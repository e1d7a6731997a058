"Times the training attention forward (C3 geometry: 32 streams x 8 heads, T = M = 512, dropout 0.1) through the C ABI with CUDA events."
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmusicgeneration_b200 import _lib
from deepmusicgeneration_b200._lib import check
B, T, H, M = int(os.environ.get('B', 32)), 512, 8, 512
HD = H * 64
lib = _lib.load()
g = torch.Generator(device='cuda').manual_seed(0)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
NB = 6   # rotate buffers: 6 x (50 + 34) MB > L2
qkv = [(torch.randn(B * T, 3 * HD, device='cuda', generator=g) * 0.8).bfloat16() for _ in range(NB)]
kvm = [(torch.randn(B * M, 2 * HD, device='cuda', generator=g) * 0.8).bfloat16() for _ in range(NB)]
rk = (torch.randn(M + T, HD, device='cuda', generator=g) * 0.8).bfloat16()
u = torch.randn(HD, device='cuda', generator=g) * 0.3
v = torch.randn(HD, device='cuda', generator=g) * 0.3
out = torch.zeros(B * T, HD, device='cuda', dtype=torch.bfloat16)
lse = torch.zeros(B, H, T, device='cuda')
def run(i, pdrop=0.1):
    check(lib.dmg_attn_train_fwd(p(qkv[i % NB]), 3 * HD, p(kvm[i % NB]), 2 * HD, p(rk), p(u), p(v), p(out), p(lse), B, T, H, M, M, 1, 1, pdrop, 99, st), 'fwd')
for pdrop in (0.1, 0.0):
    for i in range(5): run(i, pdrop)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 30
    e0.record()
    for i in range(R): run(i, pdrop)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / R * 1e3
    dense = 3 * 2 * B * T * HD * (M + T)
    print(f'attn_train_fwd p={pdrop} tc={"off" if os.environ.get("DMG_ATTN_FWD_MMA_SYNC") else "on"}: {us:.1f} us  {dense / us / 1e6:.0f} TFLOP/s dense-count')
