"""Helpers shared by the parity tests: load a golden fixture and rebuild its inputs."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

VI_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN)
                  if f.endswith('.npz') and (f.startswith('syn_') or f.startswith('vischeme_')))


# tests/golden/extra/: the reference's default two-cohort grid (582 components) and 4 / 6 cohorts
EXTRA_CASES = sorted(f[:-4] for f in os.listdir(os.path.join(GOLDEN, 'extra')) if f.endswith('.npz')) \
    if os.path.isdir(os.path.join(GOLDEN, 'extra')) else []


def load_case(name):
    sub = 'extra' if name.startswith('xtr_') else ''
    return dict(np.load(os.path.join(GOLDEN, sub, name + '.npz'), allow_pickle=False))


def build_ld(fx, block_cls, bdm_cls):
    """Rebuild the per-cohort LD operators of a fixture with the given classes."""
    P = fx['betas'].shape[0]
    out = []
    for p in range(P):
        pre = 'ld%d_' % p
        blocks = []
        for b in range(int(fx[pre + 'nblocks'])):
            u = fx[pre + 'u%d' % b]
            blocks.append(block_cls(u=u, s=fx[pre + 's%d' % b], v=u.T.copy(),
                                    D=fx[pre + 'D%d' % b]))
        out.append(bdm_cls(blocks, perm=fx[pre + 'perm'], missing=fx[pre + 'missing']))
    return out


def vi_kwargs(fx):
    return dict(marginal_effects=fx['betas'], std_errs=fx['std_errs'],
                mixture_covs=list(fx['mixture_covs']), annotations=fx['annotations'],
                scaled=bool(fx['scaled']), scale_se=bool(fx['scale_se']),
                gwas_N=fx['gwas_n'], init_hg=fx['init_hg'], num_its=int(fx['num_its']),
                checkpoint=False, checkpoint_freq=-1, output='unused')
