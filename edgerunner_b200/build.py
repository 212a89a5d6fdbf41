"""In-tree build of the sm_100a CUDA library (no JIT cache).  The .so is git-ignored but NOT gpurun-ignored: it is built here
(`__graft_entry__.build()` / this module) and travels to the GPU box with the working-tree snapshot; `_lib.load()` never builds.

    python -m edgerunner_b200.build [--force]      # or: from edgerunner_b200.build import build; build()
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libedgerunner_b200.so')
SOURCES = ['engine.cu', 'decode_kernel.cu', 'gemm.cu', 'gemm_tcgen05.cu', 'attention.cu', 'attention_tcgen05.cu', 'elementwise.cu', 'dit.cu', 'optim.cu', 'backward.cu', 'attention_bwd_mma.cu', 'train.cu', 'meto.cpp', 'meto_encode.cpp', 'mesh_clean.cpp']
HEADERS = ['common.cuh', 'decode_kernel.h', 'decode_partition.h', 'kernels.h', 'engine_internal.h', os.path.join('..', '..', 'include', 'edgerunner_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    return 'nvcc'


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + '.o')
        cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError('nvcc failed on ' + src)
        objs.append(obj)
    r = subprocess.run([_nvcc(), '-shared', '-o', LIB] + objs, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
