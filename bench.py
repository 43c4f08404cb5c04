#!/usr/bin/env python
"""Benchmark of the `vilma fit` hot path: SNP-updates/s on BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (configs[1]): synthetic single cohort, 1.2M SNPs (1 % without LD) in 1,700 dense
LD blocks (log-normal sizes, CV 0.6), default mixture grid (-K 12 -> 14 components).
A "step" is one outer iteration of MultiPopVI.optimize() (reference
variational_inference.py:361-389); metric = M x (line-search trials executed) / time,
as defined in SURVEY.md section 8(d).  The problem size is fixed as N grows (strong scaling).

`value`  : state resident in HBM, timed with CUDA events on the launching stream.
`e2e`    : the public call -- MultiPopVI.optimize(checkpoint) with HOST parameter arrays:
           upload, K iterations, download of the fitted parameters, all inside the region.
`roofline`: the LD mat-vec kernel, timed per launch with CUDA events inside the region.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (numba + NumPy, installed in
           oracle/_ref by oracle/build_ref.py; `kind: "reference"`) on a bounded sample of the same
           workload on all host cores; the NumPy oracle port (`kind: "port"`) only if numba or
           oracle/_ref is missing, with the reason on stderr and in `cpu_baseline.note`.
`workloads`: after the headline, the other BASELINE.json configurations on the same build --
           configs[2] (c3) and configs[4] (c5, with a checkpoint-every-5 + resume leg) at N = 1,
           configs[3] (c4: 6M SNPs) at N = 8 -- each with value / roofline / convergence / cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_TOTAL = 1_200_000
N_BLOCKS = 1_700
MISSING_FRAC = 0.01
K_GRID = 12
N_GWAS = 3e5
INIT_HG = 0.3
SAMPLE_BLOCKS = int(os.environ.get('BENCH_SAMPLE_BLOCKS', '170'))  # CPU sample: 10 % of the blocks (~120k SNPs)
SAMPLE_BLOCKS_MULTI = 12    # multi-cohort workloads: the reference holds three [K,P,P,M] arrays
T_START = time.time()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------
# problem construction
# --------------------------------------------------------------------------------------
def layout(M_total=M_TOTAL, n_blocks=N_BLOCKS):
    from vilma_b200 import synth
    M_ld = int(round(M_total * (1 - MISSING_FRAC)))
    n = synth.block_sizes(M_ld, n_blocks)
    starts = np.concatenate([[0], np.cumsum(n)])
    return M_ld, n, starts


def build_gpu_problem(comm, device, M_total=M_TOTAL, n_blocks=N_BLOCKS, num_its=1000):
    """Generate this rank's LD shard in HBM and build the MultiPopVI over it."""
    import torch
    from vilma_b200 import synth
    from vilma_b200.engine import DeviceContext, DeviceLD, dense_bytes
    from vilma_b200.variational_inference import DeviceBlockDiagonalMatrix, MultiPopVI

    dev = torch.device('cuda', device)
    M_ld, n_all, starts = layout(M_total, n_blocks)
    mine = synth.assign_blocks(n_all, comm.world)[comm.rank]
    # SNPs without LD are dealt round-robin
    miss_all = np.arange(M_ld, M_total, dtype=np.int64)
    miss_mine = miss_all[comm.rank::comm.world]
    snps_ld = np.concatenate([np.arange(starts[b], starts[b + 1]) for b in mine])
    snps = np.sort(np.concatenate([snps_ld, miss_mine]))
    g2l = np.full(M_total, -1, dtype=np.int64)
    g2l[snps] = np.arange(len(snps))

    t0 = time.time()
    # pass 1: standard errors (global sum of se^-2 sets the ridge prior, reference :246-247)
    se_blocks = {}
    inv_se2 = 0.0
    for b in mine:
        se_blocks[b] = synth.block_se(int(n_all[b]), 42, b, N_GWAS, dev)
        inv_se2 += float((se_blocks[b] ** -2).sum())
    inv_se2 = float(comm.sum(np.array([inv_se2]))[0]) + (M_total - M_ld) * 1.0
    prior = 2 * N_GWAS * INIT_HG / inv_se2

    # pass 2: LD blocks straight into the library's HBM store + per-SNP set-up values
    ctx = DeviceContext(device)
    ld = DeviceLD(ctx, len(snps), n=n_all[mine], rank=-np.ones(len(mine), dtype=np.int64))
    beta_hat = np.zeros(M_total)
    se = np.ones(M_total)
    adj = np.zeros(M_total)
    inv_betas = np.zeros(M_total)
    chi = 0.0
    for j, b in enumerate(mine):
        nb = int(n_all[b])
        r, bh = synth.make_block_sumstats(nb, 42, b, se_blocks[b], M_total, dev)
        ld.set_dense(j, r)
        c, a, ib = synth.precompute_block_full_rank(r, bh, se_blocks[b], prior)
        chi += c
        sl = slice(starts[b], starts[b + 1])
        beta_hat[sl] = bh.cpu().numpy()
        se[sl] = se_blocks[b].cpu().numpy()
        adj[sl] = a.cpu().numpy()
        inv_betas[sl] = ib.cpu().numpy()
        del r, bh
    ld.finalize(g2l[snps_ld])
    torch.cuda.synchronize()
    del se_blocks
    torch.cuda.empty_cache()
    # global per-SNP vectors (each rank filled only its own entries; the rest are zero)
    if comm.world > 1:
        beta_hat, adj, inv_betas = (comm.sum(x) for x in (beta_hat, adj, inv_betas))
        se = comm.sum(se - 1.0) + 1.0
    chi = float(comm.sum(np.array([chi]))[0])
    ld_diags = np.zeros(M_total)
    ld_diags[:M_ld] = 1.0
    covs = synth.mixture_grid_single(beta_hat, se, K_GRID)
    pre = dict(ld_diags=ld_diags[None], adj_marginal_effects=adj[None], chi_stat=np.array([chi]),
               ld_ranks=np.array([float(M_ld)]), inverse_betas=inv_betas[None])
    vi = MultiPopVI(marginal_effects=beta_hat[None], std_errs=se[None],
                    ld_mats=[DeviceBlockDiagonalMatrix(ld, (M_total, M_total))],
                    mixture_covs=covs, annotations=np.ones((M_total, 1)), checkpoint=False,
                    checkpoint_freq=-1, output='bench', scaled=False, scale_se=False,
                    gwas_N=np.array([N_GWAS]), init_hg=np.array([INIT_HG]), num_its=num_its,
                    comm=comm, device=device, precomputed=pre, local_snps=snps, context=ctx)
    info = dict(M=M_total, M_ld=M_ld, blocks=int(n_blocks), K=len(covs), P=1,
                ld_bytes_total=int(sum(dense_bytes(int(v)) for v in n_all)),
                ld_bytes_rank=int(ld.bytes), setup_s=time.time() - t0,
                n_max=int(n_all.max()))
    return vi, ctx, info


# Other BASELINE.json configurations (not the default bench line; `--workload c3|c5`).  Same block
# layout and sumstat model; P cohorts with their own LD (seed per cohort), shared causal SNPs with
# cross-cohort effect correlation 0.8.
WORKLOADS = {
    'c3': dict(P=3, N=(3e5, 1e5, 5e4), n_ref=0.3, ldthresh=0.99, grid=('simple', 2),
               name='BASELINE configs[2]: synthetic 3-cohort joint fit, per-cohort low-rank LD '
                    '(reference panel of 0.3 n haplotypes, --ldthresh 0.99), -K 2 grid restricted to its '
                    'positive-definite members (the reference constructor rejects the others, '
                    'variational_inference.py:610-613): 87 components; blocks stored as read-once factors'),
    'c5': dict(P=5, N=(3e5, 1e5, 5e4, 5e4, 2e4), n_ref=2.0, ldthresh=1.0, grid=('custom', 256),
               name='BASELINE configs[4]: synthetic 5-cohort fit, dense per-cohort LD, custom grid of '
                    '256 SPD covariance matrices'),
}


def mixture_grid_multi(beta_hat, se, P, spec):
    """Covariance grid of a multi-cohort workload: the reference's _make_simple (vi_options.py:301-337,
    positive-definite members only -- its own constructor rejects the rest, :610-613) or a custom grid
    built the same way (log-spaced scales x 3 random rescalings) as --load-checkpoint's .pkl allows."""
    from vilma_b200 import vi_options
    mins, maxes = vi_options._grid_range(beta_hat, se, False)
    np.random.seed(42)
    kind, k = spec
    if kind == 'simple':
        covs = vi_options._make_simple(P, k, mins, maxes)
    else:
        levels = (k + 2) // 3
        diag_vals = vi_options._make_diag_vals(P, levels - 2, mins, maxes)
        covs = []
        for idx, diag in enumerate(diag_vals):
            rho = (0.0, 0.5, 0.9)[idx % 3]
            mat = np.full((P, P), rho) + (1 - rho) * np.eye(P)
            mat = mat * np.sqrt(diag)
            mat = mat.T * np.sqrt(diag)
            for _ in range(3):
                scale = np.diag(np.sqrt(np.exp(np.random.uniform(-1, 1, P))))
                covs.append(scale.dot(mat.dot(scale)))
        covs = covs[:k]
    return [c for c in covs if np.all(np.linalg.eigvalsh(c) > 0)]


def build_gpu_problem_multi(comm, device, wl, M_total=M_TOTAL, n_blocks=N_BLOCKS, num_its=1000):
    """Multi-cohort analogue of build_gpu_problem: every cohort's LD shard is generated in HBM."""
    import torch
    from vilma_b200 import synth
    from vilma_b200.engine import DeviceContext, DeviceLD, choose_storage, dense_bytes, factor_bytes
    from vilma_b200.variational_inference import DeviceBlockDiagonalMatrix, MultiPopVI

    dev = torch.device('cuda', device)
    P, N = wl['P'], np.array(wl['N'], dtype=np.float64)
    M_ld, n_all, starts = layout(M_total, n_blocks)
    mine = synth.assign_blocks(n_all, comm.world)[comm.rank]
    miss_all = np.arange(M_ld, M_total, dtype=np.int64)
    snps_ld = np.concatenate([np.arange(starts[b], starts[b + 1]) for b in mine])
    snps = np.sort(np.concatenate([snps_ld, miss_all[comm.rank::comm.world]]))
    g2l = np.full(M_total, -1, dtype=np.int64)
    g2l[snps] = np.arange(len(snps))
    t0 = time.time()
    low_rank = wl['ldthresh'] < 1.0 or wl['n_ref'] <= 1.0

    def rank_of(n):
        return max(2, int(np.ceil(wl['n_ref'] * n))) - 1 if low_rank else n

    inv_se2 = np.zeros(P)
    for b in mine:
        for p in range(P):
            inv_se2[p] += float((synth.block_se(int(n_all[b]), 42 + p, b, N[p], dev) ** -2).sum())
    inv_se2 = comm.sum(inv_se2) + (M_total - M_ld) * 1.0
    prior = 2 * N * INIT_HG / inv_se2

    ctx = DeviceContext(device)
    kinds = [choose_storage(int(n_all[b]), rank_of(int(n_all[b]))) if low_rank else 'dense' for b in mine]
    ranks = np.array([-1 if k == 'dense' else rank_of(int(n_all[b])) for k, b in zip(kinds, mine)],
                     dtype=np.int64)
    lds = [DeviceLD(ctx, len(snps), n=n_all[mine], rank=ranks) for _ in range(P)]
    beta_hat = np.zeros((P, M_total))
    se = np.ones((P, M_total))
    adj = np.zeros((P, M_total))
    inv_betas = np.zeros((P, M_total))
    ld_diags = np.zeros((P, M_total))
    chi = np.zeros(P)
    ld_ranks = np.zeros(P)
    for j, b in enumerate(mine):
        nb = int(n_all[b])
        sl = slice(starts[b], starts[b + 1])
        beta = synth.shared_effects(nb, 42, b, P, M_total, dev)
        for p in range(P):
            se_b = synth.block_se(nb, 42 + p, b, N[p], dev)
            blk = synth.cohort_block(nb, 1000 + p, b, p, se_b, beta[p], dev, wl['n_ref'], wl['ldthresh'])
            ib = synth.ridge_start(blk, se_b, float(prior[p]))
            if kinds[j] == 'dense':
                r = blk['R'] if blk['R'] is not None else (blk['U'] * blk['s'][None, :]) @ blk['U'].T
                lds[p].set_dense(j, r.contiguous())
            else:
                if blk['rank'] != ranks[j]:
                    raise RuntimeError('block %d cohort %d: rank %d, planned %d' % (b, p, blk['rank'], ranks[j]))
                lds[p].set_factor(j, blk['U'], blk['s'])
            chi[p] += blk['chi']
            ld_ranks[p] += blk['rank']
            beta_hat[p, sl] = blk['beta_hat'].cpu().numpy()
            se[p, sl] = se_b.cpu().numpy()
            adj[p, sl] = blk['adj'].cpu().numpy()
            inv_betas[p, sl] = ib.cpu().numpy()
            ld_diags[p, sl] = blk['ld_diag'].cpu().numpy()
            del blk
    for ld in lds:
        ld.finalize(g2l[snps_ld])
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if comm.world > 1:
        beta_hat, adj, inv_betas, ld_diags = (comm.sum(x) for x in (beta_hat, adj, inv_betas, ld_diags))
        se = comm.sum(se - 1.0) + 1.0
    chi, ld_ranks = comm.sum(chi), comm.sum(ld_ranks)
    covs = mixture_grid_multi(beta_hat, se, P, wl['grid'])
    pre = dict(ld_diags=ld_diags, adj_marginal_effects=adj, chi_stat=chi, ld_ranks=ld_ranks,
               inverse_betas=inv_betas)
    vi = MultiPopVI(marginal_effects=beta_hat, std_errs=se,
                    ld_mats=[DeviceBlockDiagonalMatrix(ld, (M_total, M_total)) for ld in lds],
                    mixture_covs=covs, annotations=np.ones((M_total, 1)), checkpoint=False,
                    checkpoint_freq=-1, output='bench', scaled=False, scale_se=False,
                    gwas_N=N, init_hg=np.full(P, INIT_HG), num_its=num_its,
                    comm=comm, device=device, precomputed=pre, local_snps=snps, context=ctx)
    per_block = [dense_bytes(int(n)) if choose_storage(int(n), rank_of(int(n))) == 'dense' or not low_rank
                 else factor_bytes(int(n), rank_of(int(n))) for n in n_all]
    info = dict(M=M_total, M_ld=M_ld, blocks=int(n_blocks), K=len(covs), P=P,
                ld_bytes_total=int(P * sum(per_block)), ld_bytes_rank=int(sum(ld.bytes for ld in lds)),
                setup_s=time.time() - t0, n_max=int(n_all.max()),
                ld_store='%d of %d blocks symmetric-packed dense, the rest as factors U sqrt(s) read once per mat-vec' % (
                    sum(1 for n in n_all if not low_rank or choose_storage(int(n), rank_of(int(n))) == 'dense'),
                    len(n_all)),
                rank_total=float(ld_ranks.sum()), name=wl['name'])
    return vi, ctx, info


def cpu_classes(kind='port'):
    """(LowRank block class, block-diagonal class, VI class, extra VI kwargs, kind, note) of the CPU arm:
    the unmodified reference from oracle/_ref when it imports here, else the NumPy oracle port."""
    note = ''
    if kind == 'reference':
        try:
            from oracle import ref_loader
            ref = ref_loader.import_reference()
            return (ref.matrix_structures.LowRankMatrix, ref.matrix_structures.BlockDiagonalMatrix,
                    ref.variational_inference.MultiPopVI,
                    dict(checkpoint=False, checkpoint_freq=-1, output='bench_ref'), 'reference', note)
        except Exception as exc:
            note = 'reference unavailable, timing the NumPy port instead: %s' % (exc,)
            log('[cpu arm] ' + note)
    from oracle.ld_np import BlockDiagonalLD, LowRankBlock
    from oracle.vi_np import OracleVI
    return LowRankBlock, BlockDiagonalLD, OracleVI, {}, 'port', note


class TrialCounter:
    """Counts _update_beta line-search trials and LD mat-vecs of a reference / oracle VI object the way
    tests/golden/make_golden.py's Recorder does (objective evaluations inside _update_beta, minus the
    one that evaluates the incoming state)."""

    def __init__(self, vi):
        self.trials = 0
        self.matvec = 0
        self.iterations = 0
        me = self
        orig_step = vi._optimize_step

        def step(*a, **k):
            me.iterations += 1
            return orig_step(*a, **k)
        vi._optimize_step = step
        if hasattr(vi, 'counters'):          # the port counts trials / mat-vecs for itself
            self._vi = vi
            return
        self._vi = None
        orig_obj_fn, orig_update = vi._beta_objective, vi._update_beta

        def beta_obj(params):
            me.trials += 1
            return orig_obj_fn(params)

        def update_beta(vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
            if orig_obj is None:
                me.trials -= 1
            return orig_update(vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr)
        vi._beta_objective, vi._update_beta = beta_obj, update_beta
        for ld in vi.ld_mats:
            orig_dot = ld.dot

            def dot(x, _f=orig_dot):
                me.matvec += 1
                return _f(x)
            ld.dot = dot

    def read(self):
        if self._vi is not None:
            return self._vi.counters['trials'], self._vi.counters['matvec']
        return self.trials, self.matvec


def build_cpu_sample(n_blocks=SAMPLE_BLOCKS, kind='port', sizes=None):
    """The bounded CPU sample: the first `n_blocks` blocks of the same generator (or blocks of the
    given `sizes`), as reference / oracle objects (LowRankMatrix(X=...) does the reference's
    eigendecomposition, matrix_structures.py:15-28)."""
    import torch
    from vilma_b200 import synth
    LowRankBlock, BlockDiagonalLD, OracleVI, vi_kw, kind, _ = cpu_classes(kind)

    if sizes is None:
        M_ld_full, n_all, _ = layout()
        n = n_all[:n_blocks]
        m_norm = M_TOTAL
    else:
        n = np.asarray(sizes, dtype=np.int64)
        n_blocks = len(n)
        m_norm = int(round(n.sum() / (1 - MISSING_FRAC)))
    M_ld = int(n.sum())
    n_miss = int(round(M_ld * MISSING_FRAC / (1 - MISSING_FRAC)))
    M = M_ld + n_miss
    dev = torch.device('cpu')
    blocks, bh, se = [], [], []
    for b in range(n_blocks):
        s = synth.block_se(int(n[b]), 42, b, N_GWAS, dev)
        r, beta_hat = synth.make_block_sumstats(int(n[b]), 42, b, s, m_norm, dev)
        blocks.append(LowRankBlock(X=r.numpy(), t=1.0))
        bh.append(beta_hat.numpy())
        se.append(s.numpy())
    missing = np.arange(M_ld, M, dtype=np.int64)
    ld = BlockDiagonalLD(blocks, perm=np.arange(M), missing=missing)
    beta_hat = np.concatenate(bh + [np.zeros(n_miss)])
    std = np.concatenate(se + [np.ones(n_miss)])
    covs = synth.mixture_grid_single(beta_hat, std, K_GRID)
    vi = OracleVI(marginal_effects=beta_hat[None], std_errs=std[None], ld_mats=[ld],
                  mixture_covs=covs, annotations=np.ones((M, 1)), scaled=False, scale_se=False,
                  gwas_N=np.array([N_GWAS]), init_hg=np.array([INIT_HG]), num_its=1000, **vi_kw)
    return vi, M, n_blocks


def build_cpu_sample_multi(wl, n_blocks, kind='port'):
    """Bounded CPU sample of a multi-cohort workload: the first `n_blocks` blocks of the same
    generator for every cohort, as reference / oracle objects (the reference's set-up runs on them
    unchanged)."""
    import torch
    from vilma_b200 import synth
    LowRankBlock, BlockDiagonalLD, OracleVI, vi_kw, kind, _ = cpu_classes(kind)

    P, N = wl['P'], np.array(wl['N'], dtype=np.float64)
    _, n_all, _ = layout()
    n = n_all[:n_blocks]
    M_ld = int(n.sum())
    n_miss = int(round(M_ld * MISSING_FRAC / (1 - MISSING_FRAC)))
    M = M_ld + n_miss
    dev = torch.device('cpu')
    blocks = [[] for _ in range(P)]
    beta_hat = np.zeros((P, M))
    se = np.ones((P, M))
    pos = 0
    for b in range(n_blocks):
        nb = int(n[b])
        beta = synth.shared_effects(nb, 42, b, P, M_TOTAL, dev)
        for p in range(P):
            se_b = synth.block_se(nb, 42 + p, b, N[p], dev)
            blk = synth.cohort_block(nb, 1000 + p, b, p, se_b, beta[p], dev, wl['n_ref'], wl['ldthresh'])
            if blk['R'] is not None:
                blocks[p].append(LowRankBlock(X=blk['R'].numpy(), t=1.0))
            else:
                u = blk['U'].numpy()
                blocks[p].append(LowRankBlock(u=u, s=blk['s'].numpy(), v=u.T.copy(), D=np.zeros(nb),
                                              t=wl['ldthresh']))
            beta_hat[p, pos:pos + nb] = blk['beta_hat'].numpy()
            se[p, pos:pos + nb] = se_b.numpy()
        pos += nb
    missing = np.arange(M_ld, M, dtype=np.int64)
    lds = [BlockDiagonalLD(blocks[p], perm=np.arange(M), missing=missing) for p in range(P)]
    covs = mixture_grid_multi(beta_hat, se, P, wl['grid'])
    vi = OracleVI(marginal_effects=beta_hat, std_errs=se, ld_mats=lds, mixture_covs=covs,
                  annotations=np.ones((M, 1)), scaled=False, scale_se=False, gwas_N=N,
                  init_hg=np.full(P, INIT_HG), num_its=1000, **vi_kw)
    return vi, M, n_blocks


def _all_cores():
    """Every host core for BLAS and numba (torchrun exports OMP_NUM_THREADS=1, which would otherwise
    silently make the baseline single-threaded).  Returns (cores, restore callable)."""
    cores = os.cpu_count() or 1
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMBA_NUM_THREADS'):
        os.environ[var] = str(cores)
    limiter = None
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=cores)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    try:
        import numba
        numba.set_num_threads(min(cores, numba.config.NUMBA_NUM_THREADS))
    except Exception:
        pass
    return cores, (limiter.restore_original_limits if limiter is not None else (lambda: None))


def time_cpu(steps, warmup, workload='c2', kind='reference', converge_20k=False):
    """The reference's CPU fit (or the port) on the bounded sample: returns
    (value, seconds, trials, M, info)."""
    cores, restore = _all_cores()
    try:
        return _time_cpu(steps, warmup, cores, workload, kind, converge_20k)
    finally:
        restore()


CONV20K_SIZES = [500] * 40      # M = 20,000 in 40 dense blocks of 500, K = 14


def cpu_convergence_20k(kind):
    """Time to the reference's own stopping rule on the M = 20k problem BASELINE.md timed (38.4 s on 8
    cores): the CPU arm's whole optimize() call from the seeded start."""
    vi, M, _ = build_cpu_sample(kind=kind, sizes=CONV20K_SIZES)
    vi.num_its = 2000
    cnt = TrialCounter(vi)
    np.random.seed(42)
    t0 = time.perf_counter()
    params = vi.optimize(None)
    dt = time.perf_counter() - t0
    trials, _ = cnt.read()
    pm = vi.real_posterior_mean(*params)
    return dict(seconds=dt, iterations=int(cnt.iterations), trials=int(trials), M=int(M),
                final_elbo=float(vi.elbo(params)), pm_abs_sum=float(np.abs(pm).sum()))


def _time_cpu(steps, warmup, cores, workload='c2', kind='reference', converge_20k=False):
    t_setup = time.time()
    _, _, _, _, kind_used, note = cpu_classes(kind)
    if workload in ('c2', 'c4'):         # c4: same model and grid; the CPU sample keeps C2's block sizes
        vi, M, nb = build_cpu_sample(kind=kind_used)
    else:
        vi, M, nb = build_cpu_sample_multi(WORKLOADS[workload], SAMPLE_BLOCKS_MULTI, kind=kind_used)
    cnt = TrialCounter(vi)
    np.random.seed(42)
    params = vi._initialize()
    elbo = vi.elbo(params)
    L = np.ones(5)
    running = None
    t_setup = time.time() - t_setup
    for _ in range(warmup):            # (also absorbs what is left of the numba JIT)
        params, L, elbo, running = vi._optimize_step(params, L=L, curr_elbo=elbo,
                                                     line_search_rate=2.,
                                                     running_elbo_delta=running)
        params = tuple(params)
    t0 = time.perf_counter()
    tr0, mv0 = cnt.read()
    for _ in range(steps):
        params, L, elbo, running = vi._optimize_step(params, L=L, curr_elbo=elbo,
                                                     line_search_rate=2.,
                                                     running_elbo_delta=running)
        params = tuple(params)
    dt = time.perf_counter() - t0
    tr1, mv1 = cnt.read()
    trials = tr1 - tr0
    what = ('the unmodified reference (oracle/_ref: numba njit(parallel) kernels + NumPy/OpenBLAS)'
            if kind_used == 'reference' else 'NumPy/OpenBLAS oracle port of the reference loop')
    sample = ('first %d of the %d LD blocks of the same generator (M=%d SNPs), %d warm-up + %d '
              'timed outer iterations, %d trials, %d LD mat-vecs; %s' % (nb, N_BLOCKS, M, warmup, steps,
                                                                         trials, mv1 - mv0, what))
    extra = dict(cores=cores, sample=sample, setup_s=t_setup, kind=kind_used, note=note)
    if converge_20k:
        try:
            extra['convergence_20k'] = cpu_convergence_20k(kind_used)
        except Exception as exc:
            log('cpu convergence_20k failed: %r' % (exc,))
    return M * trials / dt, dt, trials, M, extra


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
_CLOCK_POLLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
bus, period = sys.argv[1], float(sys.argv[2])
try:
    h = nv.nvmlDeviceGetHandleByPciBusId(bus.encode())
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[3]))
reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
print('max', nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
while True:
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons(h)), flush=True)
    time.sleep(period)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region by a separate NVML polling
    process (one sample every 5 ms; the region of a default run is ~45 ms, too short for
    `nvidia-smi -lms`).  A polling THREAD in this process was measured to cost rank 0 ~0.4 ms of host
    time per outer iteration (GIL / driver-lock hand-offs with the launch thread), hence the process:
    it is started early, and only the samples stamped inside [mark_begin, mark_end] are used."""
    # nvmlClocksEventReason* bit masks
    REASONS = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20,
               'hw_thermal_slowdown': 0x40}

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None
        self.period = float(os.environ.get('BENCH_CLOCK_MS', '5')) * 1e-3

    def start(self):
        if self.period <= 0:
            return
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.device)
            bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            fd, self.path = tempfile.mkstemp(suffix='.clk')
            os.close(fd)
            self.proc = subprocess.Popen([sys.executable, '-c', _CLOCK_POLLER, bus, str(self.period),
                                          str(self.device)],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mask, mx, nearest = [], 0, None, None
        try:
            for line in open(self.path):
                parts = line.split()
                if len(parts) == 2 and parts[0] == 'max':
                    mx = float(parts[1])
                elif len(parts) == 3:
                    t, c, m = float(parts[0]), float(parts[1]), int(parts[2])
                    if self.t0 is not None and self.t0 <= t <= self.t1:
                        sm.append(c)
                        mask |= m
                    elif self.t0 is not None and (nearest is None or abs(t - self.t0) < nearest[0]):
                        nearest = (abs(t - self.t0), c, m)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm and nearest is not None and nearest[0] < 0.05:
            sm, mask = [nearest[1]], nearest[2]        # region shorter than one polling period
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx,
                       reasons=sorted(k for k, m in self.REASONS.items() if mask & m), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------
# arms
# --------------------------------------------------------------------------------------
def claim_stdout():
    """Route everything that writes to fd 1 (NCCL's version banner, library chatter) to stderr and
    return a private handle on the real stdout: the bench prints exactly ONE JSON line there."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def emit(real_stdout, result):
    os.write(real_stdout, (json.dumps(result) + '\n').encode())


def setup_ranks():
    import torch
    from vilma_b200.dist import SingleComm, TorchComm
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
        os.environ['NCCL_DEBUG'] = 'WARN'       # keep NCCL's version banner off stdout (one JSON line)
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; vilma_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        comm = TorchComm()
    else:
        comm = SingleComm()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    return comm, local_rank, hbm_peak, peak_src


def build_workload(workload, comm, device, snps=None, blocks=N_BLOCKS):
    """(vi, ctx, info) of one BASELINE.json configuration, LD generated in HBM on this rank's shard."""
    if snps is None:
        snps = 6_000_000 if workload == 'c4' else M_TOTAL
    if workload in ('c2', 'c4'):
        vi, ctx, info = build_gpu_problem(comm, device, M_total=snps, n_blocks=blocks)
        info['name'] = ('BASELINE configs[1]: synthetic single cohort, dense LD, -K 12' if workload == 'c2'
                        else 'BASELINE configs[3]: synthetic single cohort, 6M SNPs genome-wide, dense LD '
                             'sharded over the ranks, -K 12')
        from vilma_b200.engine import sym_nmax
        _, n_all, _ = layout(snps, blocks)
        n_packed = int((n_all <= sym_nmax()).sum())
        info['ld_store'] = ('dense fp64: %d of %d blocks symmetric-packed in one piece (lower triangle in 8-row '
                            'panels), %d blocks above %d rows %s' % (
                                n_packed, len(n_all), len(n_all) - n_packed, sym_nmax(), info.get('big_store', '')))
    else:
        vi, ctx, info = build_gpu_problem_multi(comm, device, WORKLOADS[workload], M_total=snps,
                                                n_blocks=blocks)
    info['workload'] = workload
    log('[rank %d] %s built in %.1fs: %s' % (comm.rank, workload, info['setup_s'], info))
    return vi, ctx, info


def measure(vi, ctx, info, comm, device, args, hbm_peak, peak_src, sampler=None, with_e2e=True,
            converge=0):
    """One workload on the device: warm-up, the timed region (CUDA events, max over ranks), per-kernel
    roofline, the end-to-end public call and (optionally) the run to convergence."""
    import ctypes as C
    import torch
    M, P, K = info['M'], info['P'], info['K']
    # initial parameters: host arrays in pinned memory (e2e uploads them)
    np.random.seed(42)
    t0 = time.time()
    init = vi._initialize()
    info['initialize_s'] = time.time() - t0
    log('[rank %d] _initialize %.1fs' % (comm.rank, info['initialize_s']))
    pin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in init]
    init = tuple(t.numpy() for t in pin)
    ckpt = {'vi_mu': init[0], 'vi_delta': init[1], 'hyper_delta': init[2],
            'error_scaling': np.ones(P)}

    # ---------------- device-resident run: warm-up then timed ----------------
    vi._set_state(init)
    state = vi.begin_loop(init)
    state = vi.run_loop(state, args.warmup, fresh=True)
    comm.barrier()
    torch.cuda.synchronize()
    ctx.profile(True)
    launches0 = ctx.launch_count()
    trials0, evals0 = vi.n_trials, vi.n_evals
    tm0 = (C.c_double * 4)()
    ctx.lib.vb_fit_timing(ctx.handle, tm0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.mark_begin()
    ev0.record()
    state = vi.run_loop(state, args.warmup + args.steps, fresh=True)
    ev1.record()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.mark_end()
    comm.barrier()
    ms = ev0.elapsed_time(ev1)
    ms = float(comm.max(np.array([ms]))[0])
    prof = ctx.profile_read()
    tm1 = (C.c_double * 4)()
    ctx.lib.vb_fit_timing(ctx.handle, tm1)
    log('[rank %d] native loop host time in region: enqueue %.3f ms, wait %.3f ms over %d rendezvous; region %.3f ms; speculative trials used.wasted %.6f' % (
        comm.rank, (tm1[0] - tm0[0]) * 1e3, (tm1[1] - tm0[1]) * 1e3, int(tm1[2] - tm0[2]), ms, tm1[3] - tm0[3]))
    ctx.profile(False)
    trials = vi.n_trials - trials0
    evals = vi.n_evals - evals0
    launches = ctx.launch_count() - launches0
    steps_done = state['num_its'] - args.warmup
    value = M * trials / (ms * 1e-3)

    # roofline of the dominant kernel, this rank's launches
    mv_ms, mv_n = prof['ld_matvec']
    snp_ms, snp_n = prof['snp']
    mv_avg = mv_ms / max(mv_n, 1)
    per_rank = np.zeros((comm.world, 3))
    per_rank[comm.rank] = [mv_avg, snp_ms / max(snp_n, 1), len(vi._snps)]
    per_rank = comm.sum(per_rank) if comm.world > 1 else per_rank
    launches_per_cohort = max(1, int(round(mv_n / max(evals * P, 1))))     # factor blocks: two passes
    achieved = info['ld_bytes_rank'] / P / (mv_avg * launches_per_cohort * 1e-3) / 1e9 if mv_n else 0.0
    traffic = None
    if comm.world == 1 and M == M_TOTAL and info['workload'] == 'c2':     # the ncu capture is of this exact launch
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ld_matvec_traffic.json')))['bytes_per_launch']
        except Exception:
            pass
    bytes_trial = info['ld_bytes_total'] + 16 * K * (P + 1) * M + 64 * P * M
    n_loc = len(vi._snps)
    snp_bytes = (trials * (16 * K * (P + 1) + 64 * P) + (evals - trials) * (8 * K * (P + 1) + 40 * P)) * n_loc
    snp_achieved = snp_bytes / (snp_ms * 1e-3) / 1e9 if snp_ms else 0.0
    ld_dominant = mv_ms >= snp_ms
    tile = vi.num_pops >= 3 or K >= 32
    out = {
        'value': value, 'unit': 'SNP-updates/s', 'steps': int(steps_done), 'ms_per_step': ms / max(steps_done, 1),
        'gpu_launches': int(launches),
        'config': {
            'workload': '%s; %d SNPs (1%% without LD) in %d LD blocks (lognormal sizes, CV 0.6, max '
                        'n=%d), %d mixture components, P=%d' % (info['name'], M, info['blocks'],
                                                                info['n_max'], K, P),
            'M': M, 'blocks': info['blocks'], 'K': K, 'P': P,
            'ld_store': info['ld_store'],
            'ld_bytes': info['ld_bytes_total'], 'algorithmic_bytes_per_trial': bytes_trial,
            'trials': int(trials), 'state_evaluations': int(evals),
            'trials_per_step': trials / max(steps_done, 1),
            'l2': 'inputs larger than L2: the %.1f GB LD store is re-read from HBM by every '
                  'evaluation' % (info['ld_bytes_rank'] / 1e9),
            'parallelism': 'LD blocks sharded over %d rank(s), LPT by n^2; one all-reduce of '
                           '%d doubles per evaluated state' % (comm.world, 3 * P + 3),
            'setup_s': info['setup_s'], 'initialize_s': info['initialize_s'],
        },
        # the dominant kernel by time in the region: the LD mat-vec (one launch per cohort) for
        # single-cohort fits, the per-SNP update for large K (P+1) state streams
        'roofline': {'bound': 'hbm',
                     'kernel': 'vb_ld_sym_kernel' if ld_dominant else ('vb_snp_tile_kernel' if tile else 'vb_snp3_kernel'),
                     'achieved': achieved if ld_dominant else snp_achieved,
                     'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': (achieved if ld_dominant else snp_achieved) / hbm_peak,
                     'traffic': traffic if ld_dominant else None, 'peak_source': peak_src,
                     'bytes_per_launch': info['ld_bytes_rank'] / P / launches_per_cohort if ld_dominant else snp_bytes / max(snp_n, 1),
                     'avg_launch_ms': mv_avg if ld_dominant else snp_ms / max(snp_n, 1),
                     'launches_timed': int(mv_n if ld_dominant else snp_n),
                     'share_of_step': (mv_ms if ld_dominant else snp_ms) / ms if ms else None,
                     'ld_kernel': {'achieved': achieved, 'frac': achieved / hbm_peak, 'avg_launch_ms': mv_avg,
                                   'share_of_step': mv_ms / ms if ms else None},
                     'snp_kernel': {'achieved': snp_achieved, 'frac': snp_achieved / hbm_peak,
                                    'share_of_step': snp_ms / ms if ms else None},
                     'snp_kernel_avg_ms': snp_ms / max(snp_n, 1),
                     'finish_kernel_avg_ms': prof['finish'][0] / max(prof['finish'][1], 1),
                     'bookkeeping_ms_per_step': prof['bookkeeping'][0] / max(steps_done, 1),
                     'per_rank_ld_ms': [round(float(v), 4) for v in per_rank[:, 0]],
                     'per_rank_snp_ms': [round(float(v), 4) for v in per_rank[:, 1]],
                     'per_rank_snps': [int(v) for v in per_rank[:, 2]],
                     'whole_trial_frac': (bytes_trial / comm.world) * evals / (ms * 1e-3) / 1e9 / hbm_peak},
    }

    # ---------------- end-to-end: the public call with host buffers ----------------
    if with_e2e or converge:
        # warm the page-locked host allocator: the first cudaHostAlloc of a result buffer costs ~50 ms for C2
        # and ~7 s for C5's 14.8 GB -- a one-off of the process, not of the call being timed
        vi._resident = None
        del_me = vi._download()
        del del_me
        vi._resident = None
    if with_e2e:
        vi.num_its = args.steps
        comm.barrier()
        torch.cuda.synchronize()
        tr0 = vi.n_trials
        marks = {}

        def timed(name, fn):
            def wrapper(*a, **k):
                t = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                marks[name] = marks.get(name, 0.0) + time.perf_counter() - t
                return r
            wrapper.__wrapped__ = fn
            return wrapper
        vi.begin_loop = timed('upload+first evaluation', vi.begin_loop)
        vi._upload = timed('(upload alone)', vi._upload)
        vi.run_loop = timed('iterations', vi.run_loop)
        vi._download = timed('download', vi._download)
        t0 = time.perf_counter()
        vi.optimize(ckpt)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        log('[rank %d] e2e %.1f ms: %s' % (comm.rank, e2e_s * 1e3,
                                           ', '.join('%s %.1f ms' % (k, v * 1e3) for k, v in marks.items())))
        vi.begin_loop, vi.run_loop, vi._download, vi._upload = (
            getattr(f, '__wrapped__', f) for f in (vi.begin_loop, vi.run_loop, vi._download, vi._upload))
        e2e_s = float(comm.max(np.array([e2e_s]))[0])
        e2e_trials = vi.n_trials - tr0
        e2e_steps = vi.num_its_run
        n_local = len(vi._snps)
        h2d = (K * n_local * (P + 1) + K) * 8 / max(e2e_steps, 1)
        d2h = (K * n_local * (P + 1)) * 8 / max(e2e_steps, 1) + (3 * P + 3 + 10) * 8 * 2
        out['e2e'] = {'value': M * e2e_trials / e2e_s, 'unit': 'SNP-updates/s', 'h2d_bytes_per_step': h2d,
                      'd2h_bytes_per_step': d2h, 'steps': int(e2e_steps), 'seconds': e2e_s,
                      'call': 'MultiPopVI.optimize(checkpoint) with pinned host parameter arrays; every rank '
                              'moves only its own shard across PCIe'}
    # time to ELBO convergence (the second half of BASELINE.json's metric): the same public call run
    # until the reference's own stopping rule (variational_inference.py:376-382) fires
    if converge:
        vi.num_its = converge
        vi._resident = None
        comm.barrier()
        torch.cuda.synchronize()
        tr0, rj0, ev0_ = vi.n_trials, vi.n_rejects, vi.n_evals
        if args.profile_convergence:     # per-kernel CUDA events over the whole fit (adds ~1 % of event overhead)
            ctx.profile(True)
        tmc0 = (C.c_double * 4)()
        ctx.lib.vb_fit_timing(ctx.handle, tmc0)
        t0 = time.perf_counter()
        vi.optimize(ckpt)
        torch.cuda.synchronize()
        conv_s = float(comm.max(np.array([time.perf_counter() - t0]))[0])
        cprof = ctx.profile_read() if args.profile_convergence else {}
        ctx.profile(False)
        tmc1 = (C.c_double * 4)()
        ctx.lib.vb_fit_timing(ctx.handle, tmc1)
        ctrials = int(vi.n_trials - tr0)
        out['convergence'] = {'seconds': conv_s, 'iterations': int(vi.num_its_run),
                              'converged': bool(vi.num_its_run < converge),
                              'trials': ctrials, 'state_evaluations': int(vi.n_evals - ev0_),
                              'rejected_trials': int(vi.n_rejects - rj0), 'max_iterations': converge,
                              'final_elbo': float(vi.trajectory['elbo'][-1]),
                              'snp_updates_per_s': M * ctrials / conv_s,
                              'kernel_ms': {k: round(v[0], 3) for k, v in cprof.items()},
                              'kernel_launches': {k: int(v[1]) for k, v in cprof.items()},
                              'host_enqueue_ms': (tmc1[0] - tmc0[0]) * 1e3,
                              'host_wait_ms': (tmc1[1] - tmc0[1]) * 1e3,
                              'call': 'MultiPopVI.optimize(checkpoint) from the seeded start, host '
                                      'arrays in and out'}
    return out, ckpt


def checkpoint_resume_leg(vi, ckpt, comm, its=6, freq=5):
    """BASELINE configs[4]: `--checkpoint-freq 5` and a mid-run resume from the `.npz` it wrote
    (reference variational_inference.py:362-367 and :345-352): run `its` iterations with periodic
    checkpoints, reload the last one, check that the resumed state's ELBO is the tracked one."""
    import shutil
    import torch
    need = sum(np.asarray(v).nbytes for v in ckpt.values()) * 3
    base = None
    for cand in ('/dev/shm', tempfile.gettempdir()):
        try:
            st = os.statvfs(cand)
            if st.f_bavail * st.f_frsize > need:
                base = cand
                break
        except OSError:
            pass
    if base is None:
        return {'skipped': 'no scratch space for %.1f GB of checkpoints' % (need / 1e9)}
    d = tempfile.mkdtemp(prefix='vilma_b200_ckpt_', dir=base) if comm.rank == 0 else None
    d = comm.broadcast_bytes((d or '').encode()).decode()
    out = {'checkpoint_freq': freq, 'iterations': its, 'dir': base}
    try:
        vi.checkpoint, vi.checkpoint_freq = True, freq
        vi.checkpoint_path = os.path.join(d, 'run-checkpoint')
        vi.num_its = its
        vi._resident = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vi.optimize(ckpt)
        torch.cuda.synchronize()
        out['seconds_with_checkpoints'] = time.perf_counter() - t0
        elbo_traj = list(vi.trajectory['elbo'])
        comm.barrier()
        files = sorted(f for f in os.listdir(d) if f.endswith('.npz'))
        out['files'] = files
        out['bytes_per_checkpoint'] = os.path.getsize(os.path.join(d, files[-1]))
        last_it = max(int(f.split('.')[-2]) for f in files)
        vi.checkpoint = False
        t0 = time.perf_counter()
        z = np.load(os.path.join(d, 'run-checkpoint.%d.npz' % last_it))
        vi.num_its = 2
        vi._resident = None
        vi.optimize(z)
        torch.cuda.synchronize()
        out['resume_seconds'] = time.perf_counter() - t0
        out['resumed_from_iteration'] = last_it
        # the checkpoint at iteration `it` holds the state BEFORE that iteration: its ELBO is the
        # tracked value after iteration it-1
        params = [np.asarray(z[k]) for k in ('vi_mu', 'vi_delta', 'hyper_delta')]
        vi._set_state(params)
        e_resumed = float(vi.elbo(params))
        e_tracked = float(elbo_traj[last_it - 1])
        out['elbo_at_resume'] = e_resumed
        out['elbo_tracked'] = e_tracked
        out['resume_matches'] = bool(abs(e_resumed - e_tracked) <= 1e-8 * abs(e_tracked))
    finally:
        vi.checkpoint = False
        comm.barrier()
        if comm.rank == 0:
            shutil.rmtree(d, ignore_errors=True)
    return out


def budget_left(args):
    return args.budget_s - (time.time() - T_START)


def cpu_baseline_leg(workload, converge_20k=False):
    try:
        v, dt, tr, Ms, extra = time_cpu(steps=3 if workload == 'c2' else 2, warmup=1, workload=workload,
                                        kind='reference', converge_20k=converge_20k)
        out = {'value': v, 'unit': 'SNP-updates/s', 'cores': extra['cores'], 'kind': extra['kind'],
               'sample': extra['sample'], 'seconds': dt, 'setup_s': extra['setup_s']}
        if extra.get('note'):
            out['note'] = extra['note']
        if 'convergence_20k' in extra:
            out['convergence_20k'] = extra['convergence_20k']
        return out
    except Exception as exc:                 # the GPU numbers stand without it
        log('cpu_baseline failed: %r' % (exc,))
        return None


def gpu_convergence_20k(comm, device):
    """The M = 20k problem of BASELINE.md section 2 through the REAL constructor (host LowRankMatrix
    objects: eigendecomposition, pseudo-inverse, ridge start -- nothing precomputed) and optimize()."""
    import torch
    from vilma_b200 import synth
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    n = np.asarray(CONV20K_SIZES, dtype=np.int64)
    M_ld = int(n.sum())
    n_miss = int(round(M_ld * MISSING_FRAC / (1 - MISSING_FRAC)))
    M = M_ld + n_miss
    dev = torch.device('cpu')
    blocks, bh, se = [], [], []
    t0 = time.perf_counter()
    for b in range(len(n)):
        s = synth.block_se(int(n[b]), 42, b, N_GWAS, dev)
        r, beta_hat = synth.make_block_sumstats(int(n[b]), 42, b, s, M, dev)
        blocks.append(LowRankMatrix(X=r.numpy(), t=1.0))
        bh.append(beta_hat.numpy())
        se.append(s.numpy())
    ld = BlockDiagonalMatrix(blocks, perm=np.arange(M), missing=np.arange(M_ld, M, dtype=np.int64))
    beta_hat = np.concatenate(bh + [np.zeros(n_miss)])
    std = np.concatenate(se + [np.ones(n_miss)])
    covs = synth.mixture_grid_single(beta_hat, std, K_GRID)
    vi = MultiPopVI(marginal_effects=beta_hat[None], std_errs=std[None], ld_mats=[ld], mixture_covs=covs,
                    annotations=np.ones((M, 1)), checkpoint=False, checkpoint_freq=-1, output='bench20k',
                    scaled=False, scale_se=False, gwas_N=np.array([N_GWAS]), init_hg=np.array([INIT_HG]),
                    num_its=2000, comm=comm, device=device)
    setup_s = time.perf_counter() - t0
    np.random.seed(42)
    t0 = time.perf_counter()
    params = vi.optimize(None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pm = vi.real_posterior_mean(*params)
    out = dict(seconds=dt, iterations=int(vi.num_its_run), trials=int(vi.n_trials), M=int(M),
               final_elbo=float(vi.elbo(params)), pm_abs_sum=float(np.abs(pm).sum()), setup_s=setup_s,
               call='MultiPopVI(...) real constructor + optimize(None) on host LowRankMatrix blocks')
    vi.close()
    return out


def run_ours(args):
    real_stdout = claim_stdout()
    import torch
    comm, device, hbm_peak, peak_src = setup_ranks()
    world = comm.world
    sampler = ClockSampler(device)
    if comm.rank == 0:
        sampler.start()          # a separate process; it is polling long before the timed region

    vi, ctx, info = build_workload(args.workload, comm, device, snps=args.snps, blocks=args.blocks)
    log('[rank %d] pinned to CPUs %s' % (comm.rank, getattr(vi._eng, 'cpus', None)))
    head, ckpt = measure(vi, ctx, info, comm, device, args, hbm_peak, peak_src, sampler=sampler,
                         with_e2e=True, converge=args.converge)
    clocks = sampler.stop() if comm.rank == 0 else {}
    result = {
        'metric': 'CAVI SNP-updates/sec (1.2M SNPs, 1/2/4/8 B200)',
        'value': head['value'], 'unit': 'SNP-updates/s', 'n_gpus': world, 'steps': head['steps'],
        'warmup': args.warmup, 'ms_per_step': head['ms_per_step'],
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': head['config'],
        'clocks': {k: clocks.get(k) for k in ('sm_mhz', 'sm_max_mhz', 'reasons', 'samples')},
        'e2e': head.get('e2e'), 'gpu_launches': head['gpu_launches'], 'roofline': head['roofline'],
    }
    if 'convergence' in head:
        result['convergence'] = head['convergence']
    if args.workload == 'c5' and args.checkpoint_leg:
        result['checkpoint_resume'] = checkpoint_resume_leg(vi, ckpt, comm)
    vi.close()
    ctx.close()
    del vi, ctx, ckpt
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE.json configurations on the same build -------------------
    extras = [w for w in args.extra_workloads.split(',') if w and w != 'none']
    if args.extra_workloads == 'auto':
        extras = []
        if args.workload == 'c2' and args.snps in (None, M_TOTAL):
            extras = ['c3', 'c5'] if world == 1 else (['c4'] if world == 8 else [])
    if extras:
        result['workloads'] = {}
    for w in extras:
        if budget_left(args) < 60:
            result['workloads'][w] = {'skipped': 'time budget of %d s spent' % args.budget_s}
            continue
        try:
            vi, ctx, info = build_workload(w, comm, device)
            conv = min(args.converge, 2000 if w == 'c4' else (600 if w == 'c3' else 150)) if args.converge else 0
            wl, ckpt = measure(vi, ctx, info, comm, device, args, hbm_peak, peak_src, with_e2e=False,
                               converge=conv if budget_left(args) > 120 else 0)
            if w == 'c5' and args.checkpoint_leg and budget_left(args) > 90:
                wl['checkpoint_resume'] = checkpoint_resume_leg(vi, ckpt, comm)
            vi.close()
            ctx.close()
            del vi, ctx, ckpt
            torch.cuda.empty_cache()
            wl['n_gpus'] = world
            result['workloads'][w] = wl
        except Exception as exc:
            log('workload %s failed: %r' % (w, exc))
            result['workloads'][w] = {'failed': repr(exc)}
    # the M = 20k problem of BASELINE.md section 2 to convergence, through the real constructor; the
    # reference's run of the same problem is cpu_baseline.convergence_20k (also in --impl reference)
    if world == 1 and args.workload == 'c2' and args.converge and args.snps in (None, M_TOTAL):
        try:
            result['convergence_20k'] = gpu_convergence_20k(comm, device)
        except Exception as exc:
            log('convergence_20k failed: %r' % (exc,))
    # ---------------- CPU baseline: the reference on the host cores (rank 0, N = 1 only) ---------
    if comm.rank == 0 and world == 1 and not args.no_cpu:
        result['cpu_baseline'] = cpu_baseline_leg(args.workload, converge_20k=args.workload == 'c2' and bool(args.converge)
                                                  and budget_left(args) > 330)
        for w, wl in (result.get('workloads') or {}).items():
            if 'value' in wl and budget_left(args) > 100:
                wl['cpu_baseline'] = cpu_baseline_leg(w)

    ref20 = (result.get('cpu_baseline') or {}).get('convergence_20k')
    if ref20 and 'convergence_20k' in result:
        result['convergence_20k']['same_iterations_and_trials_as_cpu'] = bool(
            ref20['iterations'] == result['convergence_20k']['iterations']
            and ref20['trials'] == result['convergence_20k']['trials'])
    if comm.rank == 0:
        emit(real_stdout, result)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    real_stdout = claim_stdout()
    v, dt, trials, M, extra = time_cpu(steps=args.steps, warmup=args.warmup, workload=args.workload,
                                       kind='reference', converge_20k=bool(args.converge) and args.workload == 'c2')
    cpu = {'value': v, 'unit': 'SNP-updates/s', 'cores': extra['cores'], 'kind': extra['kind'],
           'sample': extra['sample'], 'setup_s': extra['setup_s']}
    if extra.get('note'):
        cpu['note'] = extra['note']
    if 'convergence_20k' in extra:
        cpu['convergence_20k'] = extra['convergence_20k']
    result = {
        'impl': 'reference',
        'metric': 'CAVI SNP-updates/sec (1.2M SNPs, 1/2/4/8 B200)',
        'value': v, 'unit': 'SNP-updates/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3 / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': ('BASELINE configs[1]' if args.workload in ('c2', 'c4') else WORKLOADS[args.workload]['name'])
                               + ' (bounded sample): ' + extra['sample']},
        'cpu_baseline': cpu,
        'e2e': {'value': v, 'unit': 'SNP-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(real_stdout, result)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--snps', type=int, default=None)
    ap.add_argument('--blocks', type=int, default=N_BLOCKS)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--converge', type=int, default=2000, metavar='MAX_ITS',
                    help='also run the fit to convergence (at most MAX_ITS iterations) and report the time')
    ap.add_argument('--workload', default='c2', choices=['c2', 'c3', 'c4', 'c5'],
                    help='c2 = BASELINE configs[1] (the bench line of record); c3 / c5 = configs[2] / configs[4]; '
                         'c4 = configs[3]: 6M SNPs, 170 GB of LD -- needs --gpus 8 (21 GB per rank)')
    ap.add_argument('--extra-workloads', default='auto',
                    help="other BASELINE.json configurations measured after the headline into `workloads`: "
                         "'auto' (c3,c5 at N=1; c4 at N=8), 'none', or a comma list")
    ap.add_argument('--no-checkpoint-leg', dest='checkpoint_leg', action='store_false',
                    help='skip the checkpoint-every-5 + resume leg of c5')
    ap.add_argument('--profile-convergence', action='store_true',
                    help='time every kernel of the run to convergence with CUDA events (diagnosis)')
    ap.add_argument('--budget-s', type=int, default=int(os.environ.get('BENCH_BUDGET_S', '540')),
                    help='wall-clock budget: optional legs are skipped once it is spent')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
