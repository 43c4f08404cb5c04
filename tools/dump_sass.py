#!/usr/bin/env python
"""Trimmed SASS listings of the hot kernels for profiles/ (cuobjdump -sass of libvilma_b200.so).

    python tools/dump_sass.py profiles/r02_sass

Writes one `<kernel>.sass` per hot kernel (function header, instruction stream without the
encoding columns) and `summary.txt` with the counts of the mnemonics that identify the design:
UBLKCP (1-D TMA bulk copies), SYNCS (mbarrier), LDS.128, DFMA / DMUL / DADD, MUFU.RCP64H, BAR."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get('VILMA_B200_LIB', os.path.join(ROOT, 'vilma_b200', 'libvilma_b200.so'))
KERNELS = {
    'vb_ld_sym_kernel': r'vb_ld_sym_kernel',
    'vb_ld_matvec_kernel': r'vb_ld_matvec_kernel',
    'vb_ld_fac_kernel': r'vb_ld_fac_kernel',
    'vb_ld_finish_sym_kernel': r'vb_ld_finish_sym_kernel',
    'vb_snp3_kernel_P1_trial_park': r'vb_snp3_kernelILi1ELi0ELb1EE',
    'vb_snp_tile_kernel_P3_trial': r'vb_snp_tile_kernelILi3ELi0EE',
    'vb_snp_tile_kernel_P5_trial': r'vb_snp_tile_kernelILi5ELi0EE',
    'vb_snp_tile_kernel_P2_trial': r'vb_snp_tile_kernelILi2ELi0EE',
}
KEYS = ['UBLKCP', 'SYNCS', 'LDS.128', 'LDS.64', 'LDS', 'LDG', 'STG', 'STS', 'DFMA', 'DMUL', 'DADD', 'MUFU',
        'SHFL', 'BAR', 'ATOM', 'RED', 'CCTL', 'LDL', 'STL']


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'sass')
    os.makedirs(out, exist_ok=True)
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r'\n\s*Function : ', txt)
    summary = []
    for name, pat in KERNELS.items():
        body = next((f for f in funcs[1:] if re.search(pat, f.split('\n', 1)[0])), None)
        if body is None:
            summary.append('%s: not found' % name)
            continue
        lines = []
        counts = collections.Counter()
        for line in body.split('\n'):
            m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\*', line)
            if not m:
                continue
            ins = m.group(2).strip()
            lines.append('%s  %s' % (m.group(1), ins))
            op = ins.split()[1] if ins.startswith('@') and len(ins.split()) > 1 else ins.split()[0]
            for k in KEYS:
                if op.startswith(k):
                    counts[k] += 1
                    break
        with open(os.path.join(out, name + '.sass'), 'w') as fh:
            fh.write('// %s\n// %s\n' % (body.split('\n', 1)[0].strip(), os.path.basename(LIB)))
            fh.write('\n'.join(lines) + '\n')
        summary.append('%s: %d instructions; %s' % (name, len(lines), ', '.join(
            '%s %d' % (k, counts[k]) for k in KEYS if counts[k])))
    with open(os.path.join(out, 'summary.txt'), 'w') as fh:
        fh.write('\n'.join(summary) + '\n')
    print('\n'.join(summary))


if __name__ == '__main__':
    main()
