#!/usr/bin/env python
"""Build a differently compiled libvilma_b200.so into variants/ and print the tile kernels' ptxas lines.

    python tools/build_variant.py pf4 -DVB_TILE_PREFETCH=4 -DVB_TILE_UNROLL_A=2

`VILMA_B200_LIB=variants/lib_pf4.so python tools/snp_bench.py ...` then measures it (variants/ is
git-ignored but travels with gpurun)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vilma_b200 import _build  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, 'variants'), exist_ok=True)
out = os.path.join(ROOT, 'variants', 'lib_%s.so' % name)
cmd = ['nvcc'] + _build.FLAGS + ['-Xptxas', '-v', '-DVB_SOURCE_HASH="%s"' % _build.source_hash()] + defs + \
    ['-o', out, os.path.join(_build.CSRC, 'vilma_b200.cu')]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode:
    sys.stderr.write(res.stderr)
    sys.exit(1)
lines = res.stderr.split('\n')
for i, l in enumerate(lines):
    m = re.search(r"Compiling entry function '_Z18vb_snp_tile_kernelILi(\d)ELi(\d)E", l)
    if m and i + 2 < len(lines):
        print('tile P=%s mode=%s: %s' % (m.group(1), m.group(2), lines[i + 2].replace('ptxas info    : ', '')))
print(out)
