#!/usr/bin/env python
"""Micro-benchmark of the LD mat-vec kernels on C2-shaped blocks (random symmetric data).

    VILMA_B200_LIB=/path/to/variant.so python tools/ld_bench.py [--frac 0.5] [--sym 1] [--reps 20]

Prints GB/s (algorithmic bytes / CUDA-event time per launch).  Used to compare kernel variants
cheaply; bench.py remains the number of record.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frac', type=float, default=1.0, help='fraction of the 1700 C2 blocks')
    ap.add_argument('--sym', type=int, default=1)
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    import torch
    from vilma_b200 import synth
    from vilma_b200.engine import DeviceContext, DeviceLD, set_option
    n_all = synth.block_sizes(1_188_000, 1700)
    nb = max(1, int(round(a.frac * len(n_all))))
    # spread the subset over the size distribution
    n = n_all[np.linspace(0, len(n_all) - 1, nb).astype(int)]
    M = int(n.sum())
    ctx = DeviceContext(0)
    set_option('ld_symmetric', a.sym)
    ld = DeviceLD(ctx, M, n=n, rank=-np.ones(nb, dtype=np.int64))
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1)
    for b in range(nb):
        r = torch.randn((int(n[b]), int(n[b])), generator=gen, device='cuda', dtype=torch.float64)
        ld.set_dense(b, (r + r.T).contiguous())
    ld.finalize(np.arange(M))
    x = torch.randn(M, generator=gen, device='cuda', dtype=torch.float64)
    y = torch.empty_like(x)
    for _ in range(3):
        ld.dot_device(x, y)
    ctx.profile(True)
    for _ in range(a.reps):
        ld.dot_device(x, y)
    ms, cnt = ctx.profile_read()['ld_matvec']
    avg = ms / cnt
    print('lib=%s sym=%d blocks=%d M=%d bytes=%.3f GB  %.4f ms/launch  %.1f GB/s' % (
        os.path.basename(os.environ.get('VILMA_B200_LIB', 'default')), a.sym, nb, M, ld.bytes / 1e9,
        avg, ld.bytes / avg / 1e6))


if __name__ == '__main__':
    main()
