"""Builds libdmg_b200.so (hand-written sm_100a CUDA + the C ABI of include/dmg_b200.h) in-tree with nvcc.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libdmg_b200.so')
SOURCES = ['gemm.cu', 'elementwise.cu', 'attention.cu', 'attention_decode2.cu', 'sampling.cu', 'model.cu']
HEADERS = ['common.cuh', 'kernels.cuh', 'sampling.cuh', os.path.join('..', '..', 'include', 'dmg_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def source_digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    stamp = LIB + '.digest'
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == source_digest()


def build(force=False, verbose=False):
    "Compile every CUDA source for sm_100a into libdmg_b200.so; no-op when the sources have not changed."
    if not force and is_current():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(LIB + '.digest', 'w') as fh:
        fh.write(source_digest())
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
