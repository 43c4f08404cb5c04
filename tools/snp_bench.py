#!/usr/bin/env python
"""Micro-benchmark of the fused per-SNP kernels (csrc/snp_kernels.cuh) at arbitrary (P, K, M).

    python tools/snp_bench.py --cases 1x14,2x582,3x123,5x256 [--snps 1200000] [--opt name=value ...]

Each cohort gets a single 64-SNP LD block (every other SNP is "without LD"), so a beta trial is the
per-SNP kernel over all M SNPs plus a negligible mat-vec: the CUDA-event time of the kernel
(vb_ctx_profile) against its HBM floor 16 K (P+1) M + 64 P M bytes.  State and grid are random
(SPD mixture covariances); numbers are throughput only -- parity lives in tests/.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def spd_grid(K, P, rng):
    covs = []
    for k in range(K):
        scale = 10.0 ** rng.uniform(-7, -3)
        a = rng.normal(size=(P, P))
        c = a @ a.T + 0.1 * np.eye(P)
        d = np.sqrt(np.diag(c))
        covs.append(scale * c / d[:, None] / d[None, :])
    return covs


def run_case(ctx, P, K, M, reps, A=1, fuse=0):
    import torch
    from vilma_b200.engine import CudaEngine, DeviceLD
    rng = np.random.default_rng(7)
    dev = torch.device('cuda', ctx.device)
    n = 64
    lds = []
    for p in range(P):
        r = rng.normal(size=(n, 2 * n))
        r = np.corrcoef(r)
        lds.append(DeviceLD(ctx, M, [dict(n=n, kind='dense', R=r)], np.arange(n, dtype=np.int64)))
    covs = spd_grid(K, P, rng)
    prec = np.stack([np.linalg.inv(c) for c in covs])
    log_det = np.array([np.linalg.slogdet(c)[1] for c in covs])
    se = 10.0 ** rng.uniform(-3, -2, size=(P, M))
    adj = rng.normal(size=(P, M)) / se
    sld = np.zeros((P, M))
    sld[:, :n] = 1.0
    sld = sld / se ** 2
    ann = rng.integers(0, A, size=M).astype(np.int32)
    eng = CudaEngine(ctx, lds, K=K, P=P, M=M, A=A, adj=adj, se=se, sld=sld,
                     scalings=np.ones((P, M)), annotations=ann, mixture_prec=prec, log_det=log_det)
    hyper = np.full((A, K), 1.0 / K)
    eng.set_hyper(hyper)
    eng.set_delta_grad(rng.normal(size=(A, K - 1)))
    eng.set_tau(np.ones(P))
    if fuse:
        from vilma_b200 import _lib
        _lib.check(ctx.lib.vb_fit_set_fusion(ctx.handle, fuse))
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    mu = 1e-3 * torch.randn((K, P, M), generator=gen, device=dev, dtype=torch.float64)
    dl = torch.softmax(torch.randn((M, K), generator=gen, device=dev, dtype=torch.float64), dim=1)
    eng.set_params_device(mu, dl.contiguous())
    del mu, dl
    torch.cuda.empty_cache()
    eng.eval()
    out = {}
    for name, fn in (('trial', lambda: eng.beta_trial(0.5)), ('refresh', eng.refresh_delta),
                     ('eval', eng.eval)):
        for _ in range(2):
            fn()
            if name != 'eval':
                eng.accept()
        ctx.sync()
        ctx.profile(True)
        for _ in range(reps):
            fn()
            if name != 'eval':
                eng.accept()
        ms, cnt = ctx.profile_read()['snp']
        ctx.profile(False)
        out[name] = ms / max(cnt, 1)
    stats = eng.eval().cpu().numpy()[:3 * P + 3]
    eng.close()
    for ld in lds:
        ld.close()
    torch.cuda.empty_cache()
    floor_trial = (16 * K * (P + 1) + 64 * P) * M
    floor_refresh = (8 * K * (P + 1) + 40 * P) * M
    return out, floor_trial, floor_refresh, stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', default='1x14,2x582,3x123,5x256')
    ap.add_argument('--snps', type=int, default=1_200_000)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--ann', type=int, default=1)
    ap.add_argument('--fuse', type=int, default=0, help='vb_fit_set_fusion value (annotation sums ride along)')
    ap.add_argument('--opt', action='append', default=[], help='vb_set_option name=value')
    ap.add_argument('--peak', type=float, default=6553.0)
    a = ap.parse_args()
    from vilma_b200.engine import DeviceContext, set_option
    ctx = DeviceContext(0)
    for o in a.opt:
        k, v = o.split('=')
        set_option(k, int(v))
    for case in a.cases.split(','):
        P, K = (int(t) for t in case.split('x'))
        M = a.snps
        # keep 2 x (mu + delta) under ~60 GB
        while 16 * K * (P + 1) * M > 60e9:
            M //= 2
        out, ft, fr, stats = run_case(ctx, P, K, M, a.reps, a.ann, a.fuse)
        print('fuse=%d ' % a.fuse + 'opts=%s P=%d K=%d M=%d A=%d  trial %.3f ms (floor %.3f, %.2f of HBM peak)  refresh %.3f ms '
              '(floor %.3f, %.2f)  eval %.3f ms  checksum %.12e' % (
                  ','.join(a.opt) or '-', P, K, M, a.ann, out['trial'], ft / a.peak / 1e6,
                  ft / a.peak / 1e6 / out['trial'], out['refresh'], fr / a.peak / 1e6,
                  fr / a.peak / 1e6 / out['refresh'], out['eval'], float(np.sum(stats))), flush=True)


if __name__ == '__main__':
    main()
