"""Build libvilma_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The library carries a SHA-256 of the sources it was compiled from (`vb_source_hash()`, also findable
in the file as the bytes after "VB_SOURCE_HASH="): `build_library()` recompiles whenever that hash
differs from the sources on disk, so a prebuilt `.so` that travelled with the tree can never be stale
silently, and `__graft_entry__.build()` asserts the match.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvilma_b200.so')
SOURCES = ['vilma_b200.cu']
HEADERS = ['vb_common.cuh', 'ld_kernels.cuh', 'snp_kernels.cuh', 'snp_tile_kernel.cuh', 'setup_kernels.cuh',
           os.path.join('..', '..', 'include', 'vilma_b200.h')]
MARKER = b'VB_SOURCE_HASH='
FLAGS = ['-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
         '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo']


def source_hash():
    """SHA-256 over the CUDA sources, headers and compile flags (hex)."""
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        path = os.path.join(CSRC, f)
        if os.path.exists(path):
            h.update(os.path.basename(f).encode() + b'\0')
            with open(path, 'rb') as fh:
                h.update(fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def library_hash(path=LIB):
    """The source hash embedded in a built library, or None."""
    try:
        with open(path, 'rb') as fh:
            data = fh.read()
    except OSError:
        return None
    i = data.find(MARKER)
    if i < 0:
        return None
    return data[i + len(MARKER):i + len(MARKER) + 64].decode('ascii', 'replace')


def _stale():
    return library_hash() != source_hash()


def build_library(force=False, verbose=False):
    """Compile the CUDA library if missing or built from other sources.  Returns its path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    tmp = LIB + '.tmp.%d' % os.getpid()
    cmd = [nvcc] + FLAGS + ['-DVB_SOURCE_HASH="%s"' % source_hash(), '-o', tmp] \
        + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libvilma_b200.so')
    os.replace(tmp, LIB)          # atomic: concurrent ranks never load a half-written file
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force=True, verbose='-v' in sys.argv))
