"""Build libvilma_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvilma_b200.so')
SOURCES = ['vilma_b200.cu']
HEADERS = ['vb_common.cuh', 'ld_kernels.cuh', 'snp_kernels.cuh', 'snp_tile_kernel.cuh',
           os.path.join('..', '..', 'include', 'vilma_b200.h')]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False):
    """Compile the CUDA library if missing or older than its sources.  Returns its path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
           '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
           '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libvilma_b200.so')
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force=True, verbose='-v' in sys.argv))
