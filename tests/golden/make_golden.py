"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  Three kinds of fixture:

* ``cli_*``    -- the reference's own end-to-end goldens
                  (tests/test.py:2161-2197 ``test_cli_fit`` -> copy_vilma_run.*;
                  example/example.sh -> copy_of_example_vilma_run.estimates.tsv;
                  example/checkpoint_example.sh -> checkpoint_example_vilma_run.*)
                  packed together with the input files they were produced from, and
                  with the output of the shimmed reference run here (which must agree
                  with the committed goldens -- asserted below).
* ``vischeme_*`` -- the reference's deterministic unit-test fixtures
                  (tests/test.py:1226-1294) run through ``optimize()`` with the full
                  ELBO / L / trial-count trajectory recorded.
* ``syn_*``    -- seeded synthetic fits (multi block x multi cohort x annotations x
                  missing SNPs x non-identity perm x low-rank blocks) covering what the
                  reference's tests do not (SURVEY.md section 4 "gaps"), again with
                  trajectories.

Everything a parity test needs (inputs, initial parameters, per-iteration
trajectory, final parameters) is inside the fixture, so the tests never touch
/root/reference.
"""
import argparse
import io
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

REF = '/root/reference'


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

def pack_ld(prefix, bdm, out):
    """Store a reference BlockDiagonalMatrix as flat arrays under `prefix`."""
    out[prefix + 'nblocks'] = np.array(len(bdm.matrices))
    out[prefix + 'perm'] = np.asarray(bdm.perm, dtype=np.int64)
    out[prefix + 'missing'] = np.asarray(bdm.missing, dtype=np.int64)
    for b, m in enumerate(bdm.matrices):
        out[prefix + 'u%d' % b] = np.array(m.u[:])
        out[prefix + 's%d' % b] = np.array(m.s)
        out[prefix + 'D%d' % b] = np.array(m.D)


class Recorder:
    """Wrap a reference MultiPopVI to capture the trajectory at full precision."""

    def __init__(self, vi):
        self.vi = vi
        self.elbo = []          # tracked ELBO *entering* each outer iteration
        self.elbo_out = []      # tracked ELBO after each outer iteration
        self.L0 = []            # L[0] after each outer iteration
        self.trials = []        # _update_beta line-search trials per outer iteration
        self.beta_calls = []    # _update_beta calls per outer iteration
        self.tau = []           # error_scaling after each outer iteration
        self.running = []
        self._n_obj = 0
        self._n_beta = 0
        orig_step = vi._optimize_step
        orig_update_beta = vi._update_beta
        orig_beta_obj = vi._beta_objective
        rec = self

        def beta_obj(params):
            rec._n_obj += 1
            return orig_beta_obj(params)

        def update_beta(vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
            rec._n_beta += 1
            # the first objective evaluation of a call with orig_obj None is not a trial
            if orig_obj is None:
                rec._n_obj -= 1
            return orig_update_beta(vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr)

        def step(params, L, curr_elbo, line_search_rate=1.25, running_elbo_delta=None):
            rec._n_obj = 0
            rec._n_beta = 0
            rec.elbo.append(float(curr_elbo))
            res = orig_step(params, L=L, curr_elbo=curr_elbo,
                            line_search_rate=line_search_rate,
                            running_elbo_delta=running_elbo_delta)
            rec.elbo_out.append(float(res[2]))
            rec.L0.append(float(res[1][0]))
            rec.trials.append(rec._n_obj)
            rec.beta_calls.append(rec._n_beta)
            rec.tau.append(np.array(vi.error_scaling, dtype=float))
            rec.running.append(float(res[3]))
            return res

        vi._optimize_step = step
        vi._update_beta = update_beta
        vi._beta_objective = beta_obj

    def dump(self, out):
        out['traj_elbo_in'] = np.array(self.elbo)
        out['traj_elbo_out'] = np.array(self.elbo_out)
        out['traj_L0'] = np.array(self.L0)
        out['traj_trials'] = np.array(self.trials, dtype=np.int64)
        out['traj_beta_calls'] = np.array(self.beta_calls, dtype=np.int64)
        out['traj_tau'] = np.array(self.tau)
        out['traj_running'] = np.array(self.running)


def run_vi_case(ref, betas, std_errs, ld_mats, mixture_covs, annotations,
                scaled, scale_se, gwas_n, init_hg, num_its, seed,
                resume_at=None):
    """Run the reference MultiPopVI and return a dict fixture."""
    vin = ref.variational_inference
    out = {}
    out['betas'] = np.array(betas, dtype=float)
    out['std_errs'] = np.array(std_errs, dtype=float)
    out['annotations'] = np.array(annotations)
    out['mixture_covs'] = np.array(mixture_covs, dtype=float)
    out['gwas_n'] = np.array(gwas_n, dtype=float)
    out['init_hg'] = np.array(init_hg, dtype=float)
    out['scaled'] = np.array(bool(scaled))
    out['scale_se'] = np.array(bool(scale_se))
    out['num_its'] = np.array(int(num_its))
    out['seed'] = np.array(int(seed))
    for p, ld in enumerate(ld_mats):
        pack_ld('ld%d_' % p, ld, out)

    def build():
        return vin.MultiPopVI(
            marginal_effects=np.array(betas, dtype=float),
            std_errs=np.array(std_errs, dtype=float),
            ld_mats=ld_mats, mixture_covs=list(mixture_covs),
            annotations=annotations, checkpoint=False, checkpoint_freq=-1,
            output='unused', scaled=scaled, scale_se=scale_se,
            gwas_N=np.array(gwas_n, dtype=float),
            init_hg=np.array(init_hg, dtype=float), num_its=num_its)

    vi = build()
    out['pre_adj_marginal_effects'] = vi.adj_marginal_effects.copy()
    out['pre_chi_stat'] = vi.chi_stat.copy()
    out['pre_ld_ranks'] = vi.ld_ranks.copy()
    out['pre_inverse_betas'] = vi.inverse_betas.copy()
    out['pre_ld_diags'] = vi.ld_diags.copy()

    # initial parameters exactly as optimize() would draw them
    np.random.seed(seed)
    init = vi._initialize()
    out['init_vi_mu'], out['init_vi_delta'], out['init_hyper_delta'] = (
        np.array(x) for x in init)
    out['init_elbo'] = np.array(float(vi.elbo(init)))
    out['init_nat_grad_vi_delta'] = np.array(vi.nat_grad_vi_delta)
    out['init_loglik'] = np.array(float(vi._log_likelihood(init)))
    out['init_beta_kl'] = np.array(float(vi._beta_KL(*init)))
    out['init_post_mean'] = vi.real_posterior_mean(*init)
    out['init_post_var'] = vi.real_posterior_variance(*init)

    # the real run (fresh object so hidden state is clean)
    vi = build()
    rec = Recorder(vi)
    np.random.seed(seed)
    params = vi.optimize(None)
    rec.dump(out)
    out['final_vi_mu'], out['final_vi_delta'], out['final_hyper_delta'] = (
        np.array(x) for x in params)
    out['final_error_scaling'] = np.array(vi.error_scaling)
    out['final_post_mean'] = vi.real_posterior_mean(*params)
    out['final_post_var'] = vi.real_posterior_variance(*params)
    out['final_vi_sigma'] = np.array(vi.vi_sigma)
    out['final_elbo_recomputed'] = np.array(float(vi.elbo(params)))

    if resume_at is not None:
        # a mid-run resume: run `resume_at` iterations, dump, resume from the dump
        vi = build()
        vi.num_its = resume_at
        np.random.seed(seed)
        mid = vi.optimize(None)
        ckpt = vi.create_dump_dict(mid)
        ckpt = {k: np.array(v) for k, v in ckpt.items()}
        for k, v in ckpt.items():
            out['resume_ckpt_' + k] = v
        vi2 = build()
        rec2 = Recorder(vi2)
        res = vi2.optimize(ckpt)
        out['resume_traj_elbo_out'] = np.array(rec2.elbo_out)
        out['resume_traj_L0'] = np.array(rec2.L0)
        out['resume_traj_trials'] = np.array(rec2.trials, dtype=np.int64)
        out['resume_final_vi_mu'] = np.array(res[0])
        out['resume_final_vi_delta'] = np.array(res[1])
        out['resume_final_hyper_delta'] = np.array(res[2])
        out['resume_final_error_scaling'] = np.array(vi2.error_scaling)
    return out


# --------------------------------------------------------------------------
# synthetic inputs (independent of the reference's sim.py code; same model)
# --------------------------------------------------------------------------

def ar1_block(n, rho, n_ref, rng):
    """Sample-correlation of n_ref AR(1) haplotype-like rows + 10% iid noise."""
    e = rng.standard_normal((n_ref, n))
    g = np.empty_like(e)
    g[:, 0] = e[:, 0]
    c = np.sqrt(1 - rho * rho)
    for j in range(1, n):
        g[:, j] = rho * g[:, j - 1] + c * e[:, j]
    g += np.sqrt(0.1) * rng.standard_normal((n_ref, n))
    g -= g.mean(axis=0)
    g /= np.sqrt((g * g).sum(axis=0))
    r = g.T @ g
    r = 0.5 * (r + r.T)
    np.fill_diagonal(r, 1.0)
    return r


def make_synthetic(ref, P, M, block_sizes, seed, ldthresh, low_rank_frac,
                   n_annot, miss_frac, shuffle, n_samp):
    ms = ref.matrix_structures
    rng = np.random.default_rng(seed)
    assert sum(block_sizes) <= M
    order = rng.permutation(M) if shuffle else np.arange(M)
    true_beta = np.zeros((P, M))
    comp = rng.choice(4, size=M, p=[0.9, 0.07, 0.02, 0.01])
    var = np.array([0.0, 1e-4, 1e-3, 1e-2])[comp]
    shared = rng.standard_normal(M)
    for p in range(P):
        own = rng.standard_normal(M)
        true_beta[p] = np.sqrt(var) * (np.sqrt(0.8) * shared + np.sqrt(0.2) * own)
    freq = rng.uniform(0.05, 0.5, size=M)
    std_errs = np.empty((P, M))
    betas = np.zeros((P, M))
    ld_mats = []
    for p in range(P):
        std_errs[p] = 1.0 / np.sqrt(n_samp[p] * 2 * freq * (1 - freq))
        prng = np.random.default_rng(1000 * seed + p)
        # SNPs missing from this cohort's LD
        n_miss = int(round(miss_frac * M))
        in_blocks = order[:sum(block_sizes)]
        drop = set(prng.choice(in_blocks, size=n_miss, replace=False).tolist()) \
            if n_miss else set()
        blocks, perm = [], []
        off = 0
        for n in block_sizes:
            idx = np.array([i for i in in_blocks[off:off + n] if i not in drop],
                           dtype=np.int64)
            off += n
            if len(idx) == 0:
                continue
            n_ref = (2 * len(idx) if low_rank_frac is None
                     else max(2, int(np.ceil(low_rank_frac * len(idx)))))
            r = ar1_block(len(idx), 0.9, n_ref, prng)
            lrm = ms.LowRankMatrix(X=r, t=ldthresh)
            blocks.append(lrm)
            perm.append(idx)
            # simulate sumstats from the *kept* low-rank LD: S R S^-1 b + S R^1/2 e
            rmat = (lrm.u * lrm.s) @ lrm.v
            s = std_errs[p, idx]
            root = (lrm.u * np.sqrt(lrm.s)) @ lrm.v
            betas[p, idx] = (s * (rmat @ (true_beta[p, idx] / s))
                             + s * (root @ prng.standard_normal(len(idx))))
        perm = np.concatenate(perm)
        missing = np.array(sorted(set(range(M)) - set(perm.tolist())), dtype=np.int64)
        full_perm = np.concatenate([perm, missing])
        ld_mats.append(ms.BlockDiagonalMatrix(blocks, perm=full_perm, missing=missing))
        # SNPs missing from LD behave like missing sumstats in the CLI (BETA 0, SE 1)
        betas[p, missing] = 0.0
        std_errs[p, missing] = 1.0
    if n_annot > 1:
        lab = rng.integers(0, n_annot, size=M)
        annotations = np.zeros((M, n_annot))
        annotations[np.arange(M), lab] = 1
    else:
        annotations = np.ones((M, 1))
    return betas, std_errs, ld_mats, annotations


def make_grid(ref, P, K, betas, std_errs, seed):
    """Mixture grid as vi_options.main builds it (vi_options.py:208-229)."""
    maxes = np.zeros(P)
    mins = np.zeros(P)
    for p in range(P):
        b = np.abs(betas[p])
        s = std_errs[p]
        psi = 1.0 / len(b)
        probs = 1. / (1. + ((1. - psi) / psi * np.sqrt(b**2 / s**2)
                            * np.exp(-0.5 * b**2 / s**2 + 0.5)))
        ebayes = np.maximum(b**2 - s**2, 1e-10)
        raw = b / (1. + s**2 / ebayes**2)
        maxes[p] = np.max(probs * raw)**2
        mins[p] = np.nanpercentile(betas[p, betas[p]**2 > 0]**2, 2.5)
    np.random.seed(seed)
    covs = ref.vi_options._make_simple(P, K, mins, maxes)
    if P > 2:
        # vi_options._make_simple emits indefinite matrices for P >= 3 (e.g. pairwise
        # correlations .99/.99/0), which MultiPopVI.__init__ rejects
        # (variational_inference.py:610-613): a P >= 3 fit needs a custom grid passed
        # via --load-checkpoint.  Keep the positive-definite members of the default grid.
        covs = [c for c in covs if np.linalg.eigvalsh(c).min() > 1e-8 * np.abs(c).max()]
    return covs


# --------------------------------------------------------------------------
# CLI goldens
# --------------------------------------------------------------------------

def read_text(path):
    with open(path, 'r') as fh:
        return fh.read()


def cli_namespace(**kw):
    ns = argparse.Namespace(
        components=12, num_its=1000, ld_schema=None, sumstats=None,
        stderrscale='1.0', annotations=None, output=None, names=None,
        extract=None, scaled=False, ldthresh=1.0, seed=42, mmap=False,
        scale_se=False, samplesizes='100e3', init_hg='0.1', trait=False,
        checkpoint_freq=-1, load_checkpoint='')
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns


def pack_outputs(prefix, out_root, out):
    npz = np.load(out_root + '.npz')
    for k in npz.files:
        out[prefix + 'npz_' + k] = npz[k]
    if os.path.exists(out_root + '.covariance.pkl'):
        with open(out_root + '.covariance.pkl', 'rb') as fh:
            out[prefix + 'covariance'] = np.array(pickle.load(fh)[0])
    out[prefix + 'estimates_tsv'] = np.array(read_text(out_root + '.estimates.tsv'))


def golden_cli_fit(ref, outdir):
    td = os.path.join(REF, 'tests', 'test_data')
    tmp = tempfile.mkdtemp()
    out = {}
    files = ['ld_manifest.tsv', 'ld_variants.tsv', 'good_sumstats_beta.tsv',
             'good_variants.tsv']
    for f in files:
        shutil.copy(os.path.join(td, f), tmp)
        out['in_' + f] = np.array(read_text(os.path.join(td, f)))
    shutil.copy(os.path.join(td, 'ld_matrix.npy'), tmp)
    out['in_ld_matrix.npy'] = np.load(os.path.join(td, 'ld_matrix.npy'))
    args = cli_namespace(
        ld_schema=os.path.join(tmp, 'ld_manifest.tsv'),
        sumstats=os.path.join(tmp, 'good_sumstats_beta.tsv'),
        output=os.path.join(tmp, 'vilma_run'), components=80, ldthresh=0.8,
        init_hg='0.2', samplesizes='10e3', names='test_cohort', scale_se=True,
        extract=os.path.join(tmp, 'good_variants.tsv'))
    out['argv'] = np.array(
        '-K 80 --ldthresh 0.8 --init-hg 0.2 --samplesizes 10e3 '
        '--names test_cohort --learn-scaling')
    ref.vi_options.main(args)
    pack_outputs('run_', os.path.join(tmp, 'vilma_run'), out)
    # the reference's own committed goldens
    pack_outputs('gold_', os.path.join(td, 'copy_vilma_run'), out)
    for k in [k for k in out if k.startswith('gold_npz_')]:
        d = np.max(np.abs(out[k] - out['run_' + k[5:]]))
        print('  cli_fit %-28s max|run-gold| = %.3e' % (k, d))
        assert d < 1e-10
    assert np.allclose(out['gold_covariance'], out['run_covariance'], rtol=1e-12, atol=0)
    np.savez_compressed(os.path.join(outdir, 'cli_fit.npz'), **out)
    shutil.rmtree(tmp)


def golden_example(ref, outdir):
    ex = os.path.join(REF, 'example')
    tmp = tempfile.mkdtemp()
    out = {}
    os.makedirs(os.path.join(tmp, 'ld_mat'))
    os.makedirs(os.path.join(tmp, 'example_data'))
    for f in ['keep_variants.txt', 'example_data/example_gwas_sumstats.txt',
              'ld_mat/example_schema.schema', 'ld_mat/example_schema_1:0.var',
              'ld_mat/example_schema_1:2.var']:
        shutil.copy(os.path.join(ex, f), os.path.join(tmp, f))
        out['in_' + f] = np.array(read_text(os.path.join(ex, f)))
    for f in ['ld_mat/example_schema_1:0.npy', 'ld_mat/example_schema_1:2.npy']:
        shutil.copy(os.path.join(ex, f), os.path.join(tmp, f))
        out['in_' + f] = np.load(os.path.join(ex, f))
    common = dict(
        ld_schema=os.path.join(tmp, 'ld_mat/example_schema.schema'),
        sumstats=os.path.join(tmp, 'example_data/example_gwas_sumstats.txt'),
        seed=42, components=81, init_hg='0.2', samplesizes='300e3', names='ukbb',
        scale_se=True, extract=os.path.join(tmp, 'keep_variants.txt'))
    out['argv'] = np.array('--seed 42 -K 81 --init-hg 0.2 --samplesizes 300e3 '
                           '--names ukbb --learn-scaling')
    ref.vi_options.main(cli_namespace(
        output=os.path.join(tmp, 'example_vilma_run'), **common))
    pack_outputs('run_', os.path.join(tmp, 'example_vilma_run'), out)
    out['gold_estimates_tsv'] = np.array(
        read_text(os.path.join(ex, 'copy_of_example_vilma_run.estimates.tsv')))
    # resume (example/checkpoint_example.sh)
    ref.vi_options.main(cli_namespace(
        output=os.path.join(tmp, 'checkpoint_example_vilma_run'),
        load_checkpoint=[os.path.join(tmp, 'example_vilma_run.npz'),
                         os.path.join(tmp, 'example_vilma_run.covariance.pkl')],
        **common))
    pack_outputs('resume_run_', os.path.join(tmp, 'checkpoint_example_vilma_run'), out)
    gold = np.load(os.path.join(ex, 'checkpoint_example_vilma_run.npz'))
    for k in gold.files:
        out['resume_gold_npz_' + k] = gold[k]
        d = np.max(np.abs(gold[k] - out['resume_run_npz_' + k]))
        print('  example resume %-16s max|run-gold| = %.3e' % (k, d))
    out['resume_gold_estimates_tsv'] = np.array(
        read_text(os.path.join(ex, 'checkpoint_example_vilma_run.estimates.tsv')))
    np.savez_compressed(os.path.join(outdir, 'cli_example.npz'), **out)
    shutil.rmtree(tmp)


def golden_cli_multi(ref, outdir):
    """A 2-cohort + annotations CLI run on the reference's own test data files.

    Exercises what test_cli_fit does not: P=2 grid (consumes np.random.uniform),
    --annotations with un-annotated variants (denylist), a stacked-eigen .npy
    schema for the second cohort, --names, --stderrscale, --checkpoint-freq.
    """
    td = os.path.join(REF, 'tests', 'test_data')
    tmp = tempfile.mkdtemp()
    out = {}
    for f in ['ld_manifest.tsv', 'ld_manifest_svd.tsv', 'ld_variants.tsv',
              'good_sumstats_beta.tsv', 'good_sumstats_flip.tsv',
              'good_variants.tsv', 'good_annotations.tsv']:
        shutil.copy(os.path.join(td, f), tmp)
        out['in_' + f] = np.array(read_text(os.path.join(td, f)))
    for f in ['ld_matrix.npy', 'ld_matrix_svd.npy']:
        shutil.copy(os.path.join(td, f), tmp)
        out['in_' + f] = np.load(os.path.join(td, f))
    args = cli_namespace(
        ld_schema=','.join([os.path.join(tmp, 'ld_manifest.tsv'),
                            os.path.join(tmp, 'ld_manifest_svd.tsv')]),
        sumstats=','.join([os.path.join(tmp, 'good_sumstats_beta.tsv'),
                           os.path.join(tmp, 'good_sumstats_flip.tsv')]),
        annotations=os.path.join(tmp, 'good_annotations.tsv'),
        output=os.path.join(tmp, 'multi_run'), components=3, ldthresh=0.9,
        init_hg='0.2,0.3', samplesizes='10e3,20e3', names='a,b',
        stderrscale='1.0,1.5', scale_se=True, seed=7, checkpoint_freq=4,
        num_its=30, extract=os.path.join(tmp, 'good_variants.tsv'))
    out['argv'] = np.array(
        '-K 3 --ldthresh 0.9 --init-hg 0.2,0.3 --samplesizes 10e3,20e3 --names a,b '
        '--stderrscale 1.0,1.5 --learn-scaling --seed 7 --checkpoint-freq 4 --num-its 30')
    ref.vi_options.main(args)
    pack_outputs('run_', os.path.join(tmp, 'multi_run'), out)
    ckpts = sorted(f for f in os.listdir(tmp) if f.startswith('multi_run-checkpoint.'))
    out['checkpoint_files'] = np.array(ckpts)
    for f in ckpts:
        z = np.load(os.path.join(tmp, f))
        for k in z.files:
            out['ckpt_%s_%s' % (f, k)] = z[k]
    np.savez_compressed(os.path.join(outdir, 'cli_multi.npz'), **out)
    shutil.rmtree(tmp)


def golden_cli_sim(ref, outdir):
    """The reference's own `vilma sim` golden (tests/test.py:2200-2246, --seed 143): its input files
    and the committed output copy_vilma_sim_run.simpop1.simgwas.tsv.  (The reference's sim path stores
    LD in HDF5 -- h5py is not installed here -- so the golden is the reference's committed file, not a
    re-run.)"""
    td = os.path.join(REF, 'tests', 'test_data')
    out = {}
    for f in ['ld_manifest.tsv', 'ld_variants.tsv', 'good_sumstats_beta.tsv', 'good_annotations.tsv']:
        out['in_' + f] = np.array(read_text(os.path.join(td, f)))
    out['in_ld_matrix.npy'] = np.load(os.path.join(td, 'ld_matrix.npy'))
    out['in_sim_weights.npy'] = np.load(os.path.join(td, 'sim_weights.npy'))
    out['weights_npz_hyper_delta'] = np.load(os.path.join(td, 'sim_weights.npz'))['hyper_delta']
    with open(os.path.join(td, 'copy_vilma_run.covariance.pkl'), 'rb') as fh:
        out['covariance'] = np.array(pickle.load(fh)[0])
    out['argv'] = np.array('--names simpop1 --seed 143')
    out['gold_simgwas_tsv'] = np.array(read_text(os.path.join(td, 'copy_vilma_sim_run.simpop1.simgwas.tsv')))
    np.savez_compressed(os.path.join(outdir, 'cli_sim.npz'), **out)


# --------------------------------------------------------------------------
# VI goldens
# --------------------------------------------------------------------------

def golden_vischeme(ref, outdir):
    """tests/test.py:1226-1294 fixtures, run through optimize()."""
    ms = ref.matrix_structures
    for linked in (True, False):
        for n_annot, scaled, scale_se in [(1, False, False), (1, True, False),
                                          (2, False, True), (2, True, True)]:
            if linked:
                betas = np.arange(100).reshape(2, 50).astype(float)
                ld = (1 + np.arange(50 * 50)).reshape(50, 50) / (50 * 50 + 1)
                ld = ld + ld.T + 5 * np.eye(50)
                d = np.diag(1 / np.sqrt(np.diag(ld)))
                ld = d @ ld @ d
            else:
                betas = np.arange(100).reshape(50, 2).T.astype(float)
                ld = np.eye(50)
            std_errs = np.array([1.] * 50 + [2.] * 50).reshape(2, 50)
            lrm = ms.LowRankMatrix(X=ld, t=1.0)
            ld_mats = [ms.BlockDiagonalMatrix([lrm]), ms.BlockDiagonalMatrix([lrm])]
            if n_annot == 2:
                ann = np.zeros((50, 2), dtype=int)
                ann[0:25, 0] = 1
                ann[25:, 1] = 1
            else:
                ann = np.ones((50, 1), dtype=int)
            name = 'vischeme_%s_a%d_s%d_t%d' % ('linked' if linked else 'unlinked',
                                                 n_annot, int(scaled), int(scale_se))
            print(name)
            fx = run_vi_case(ref, betas, std_errs, ld_mats,
                             [np.eye(2), 2 * np.eye(2)], ann, scaled, scale_se,
                             [100e3, 10e3], [0.1, 0.9], 20, seed=42)
            print('   its=%d trials=%s' % (len(fx['traj_L0']), fx['traj_trials'].tolist()))
            np.savez_compressed(os.path.join(outdir, name + '.npz'), **fx)


SYN_CASES = {
    # name: dict(P, M, blocks, K, ldthresh, low_rank_frac, n_annot, miss, shuffle, scaled, scale_se, its, resume)
    'syn_p1_dense': dict(P=1, M=420, blocks=[30, 55, 80, 41, 64, 37, 52, 48], K=12,
                         ldthresh=1.0, lrf=None, A=3, miss=0.02, shuffle=True,
                         scaled=False, scale_se=True, its=60, resume=7,
                         n=[3e4]),
    'syn_p1_scaled': dict(P=1, M=300, blocks=[60, 45, 70, 50, 64], K=6,
                          ldthresh=1.0, lrf=None, A=1, miss=0.0, shuffle=False,
                          scaled=True, scale_se=False, its=40, resume=None,
                          n=[5e4]),
    'syn_p2_lowrank': dict(P=2, M=360, blocks=[48, 66, 40, 75, 58, 62], K=2,
                           ldthresh=0.9, lrf=0.4, A=2, miss=0.03, shuffle=True,
                           scaled=False, scale_se=True, its=40, resume=5,
                           n=[4e4, 1e4]),
    'syn_p3': dict(P=3, M=240, blocks=[50, 44, 63, 38, 40], K=2,
                   ldthresh=0.99, lrf=0.5, A=1, miss=0.02, shuffle=False,
                   scaled=False, scale_se=True, its=30, resume=None,
                   n=[4e4, 2e4, 1e4]),
    'syn_p5': dict(P=5, M=150, blocks=[40, 35, 45, 25], K=1,
                   ldthresh=1.0, lrf=None, A=2, miss=0.02, shuffle=True,
                   scaled=False, scale_se=False, its=25, resume=6,
                   n=[4e4, 2e4, 1e4, 1e4, 5e3]),
}


# Further cases kept under tests/golden/extra/ (run by tests/test_gpu_extra.py): the reference's DEFAULT
# two-cohort grid (-K 12 -> 582 components: the tile kernel at 16 warps per tile) and the cohort counts
# no other fixture has (4 and 6: every P x P template instantiation then has a parity case).
EXTRA_CASES = {
    'xtr_p2_default_grid': dict(P=2, M=96, blocks=[30, 36, 28], K=12,
                                ldthresh=1.0, lrf=None, A=1, miss=0.02, shuffle=False,
                                scaled=False, scale_se=False, its=15, resume=None,
                                n=[4e4, 1e4]),
    'xtr_p4': dict(P=4, M=120, blocks=[34, 40, 44], K=1,
                   ldthresh=1.0, lrf=None, A=1, miss=0.02, shuffle=True,
                   scaled=False, scale_se=False, its=20, resume=None,
                   n=[4e4, 2e4, 1e4, 1e4]),
    'xtr_p6': dict(P=6, M=90, blocks=[28, 32, 28], K=1,
                   ldthresh=0.99, lrf=0.6, A=2, miss=0.02, shuffle=False,
                   scaled=False, scale_se=False, its=20, resume=None,
                   n=[4e4, 2e4, 1e4, 1e4, 5e3, 5e3]),
    # round 2: the shapes BASELINE.json configs[2] and [4] stress, and the multi-rank statistics vector at
    # its largest.  `grid='custom:K'` = K SPD matrices built like the reference's grid (log-spaced scales x
    # correlation levels x random rescalings), i.e. what --load-checkpoint's .pkl allows (vi_options.py:192-194).
    # C5 shape: five cohorts, 256 components (the tile kernel's 4-warp plan with 64 slots per warp)
    'xtr_p5_k256': dict(P=5, M=96, blocks=[30, 34, 28], K=None, grid='custom:256',
                        ldthresh=1.0, lrf=None, A=1, miss=0.02, shuffle=False,
                        scaled=False, scale_se=False, its=12, resume=None, slim=True,
                        n=[3e5, 1e5, 5e4, 5e4, 2e4]),
    # factor-stored LD: rank ~ n/10 (--ldthresh 0.8-like truncation), so 16 n r < 4 n (n+1) and the
    # device keeps U, s instead of the dense block (csrc/ld_kernels.cuh factor path)
    'xtr_p1_factor': dict(P=1, M=520, blocks=[120, 150, 100, 140], K=12, ldthresh=0.8, lrf=0.1,
                          A=1, miss=0.02, shuffle=True, scaled=False, scale_se=True, its=30,
                          resume=None, n=[5e4]),
    # three cohorts, A*K = 48 fused annotation sums: 3P+3+48+10 = 70 exchanged values per evaluation
    # (the multi-rank mailbox at its largest), 16 blocks so that 8 ranks all own LD
    'xtr_p3_ann48': dict(P=3, M=480, blocks=[30, 26, 34, 28, 32, 24, 36, 30, 28, 26, 32, 30, 24, 34, 28, 30],
                         K=None, grid='custom:16', ldthresh=1.0, lrf=None, A=3, miss=0.02, shuffle=True,
                         scaled=False, scale_se=True, its=25, resume=None,
                         n=[4e4, 2e4, 1e4]),
}


def custom_grid(P, K, betas, std_errs, seed):
    """K SPD P x P covariance matrices spanning the data-driven range of vi_options.py:196-229."""
    rng = np.random.default_rng(seed)
    lo = max(np.nanpercentile(betas[betas != 0]**2, 2.5), 1e-10)
    hi = max(np.max((np.abs(betas) - std_errs).clip(0)**2), 10 * lo)
    scales = np.exp(np.linspace(np.log(lo), np.log(hi), (K + 2) // 3))
    covs = []
    for idx, sc in enumerate(scales):
        rho = (0.0, 0.5, 0.9)[idx % 3]
        base = np.full((P, P), rho) + (1 - rho) * np.eye(P)
        for _ in range(3):
            d = np.sqrt(sc * np.exp(rng.uniform(-1, 1, P)))
            covs.append(base * d[:, None] * d[None, :])
    return covs[:K]


def golden_synthetic(ref, outdir, only=None, cases=None):
    for name, c in (cases or SYN_CASES).items():
        if only and name not in only:
            continue
        print(name)
        betas, std_errs, ld_mats, ann = make_synthetic(
            ref, c['P'], c['M'], c['blocks'], seed=abs(hash(name)) % 1000 if False else
            sum(map(ord, name)), ldthresh=c['ldthresh'], low_rank_frac=c['lrf'],
            n_annot=c['A'], miss_frac=c['miss'], shuffle=c['shuffle'], n_samp=c['n'])
        if c['scaled']:
            grid_b, grid_s = betas / std_errs, np.ones_like(std_errs)
        else:
            grid_b, grid_s = betas, std_errs
        if c.get('grid', '').startswith('custom:'):
            covs = custom_grid(c['P'], int(c['grid'].split(':')[1]), grid_b, grid_s, seed=11)
        else:
            covs = make_grid(ref, c['P'], c['K'], grid_b, grid_s, seed=11)
        fx = run_vi_case(ref, betas, std_errs, ld_mats, covs, ann, c['scaled'],
                         c['scale_se'], c['n'], [0.3] * c['P'], c['its'], seed=42,
                         resume_at=c['resume'])
        print('   K=%d its=%d trials=%s L0max=%.3g' % (
            len(covs), len(fx['traj_L0']), fx['traj_trials'].tolist(),
            fx['traj_L0'].max()))
        if c.get('slim'):        # keep the committed file small: [K,P,P,M] is checked on the other cases
            for k in ('final_vi_sigma',):
                fx.pop(k, None)
        np.savez_compressed(os.path.join(outdir, name + '.npz'), **fx)


def golden_cli_flags(ref, outdir):
    """`vilma fit` flag surface (vi_options.py:9-84) as a table: tests/golden/cli_fit_flags.json."""
    import json
    top = argparse.ArgumentParser(prog='vilma')
    parser = ref.vi_options.args(top.add_subparsers())
    table = {}
    for a in parser._actions:
        if not a.option_strings or a.dest == 'help':
            continue
        table[a.dest] = dict(flags=sorted(a.option_strings), default=a.default, required=bool(a.required),
                             nargs=a.nargs, type=getattr(a.type, '__name__', None), action=type(a).__name__)
    with open(os.path.join(outdir, 'cli_fit_flags.json'), 'w') as fh:
        json.dump(table, fh, indent=1, sort_keys=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', nargs='*', default=None)
    a = ap.parse_args()
    ref = import_reference()
    sel = a.only
    if not sel or 'cli' in sel:
        golden_cli_flags(ref, HERE)
        golden_cli_fit(ref, HERE)
        golden_example(ref, HERE)
        golden_cli_multi(ref, HERE)
    if not sel or 'sim' in sel:
        golden_cli_sim(ref, HERE)
    if not sel or 'vischeme' in sel:
        golden_vischeme(ref, HERE)
    if not sel or any(s.startswith('syn') for s in sel):
        golden_synthetic(ref, HERE, only=[s for s in (sel or []) if s.startswith('syn_')] or None)
    if not sel or any(s.startswith('xtr') for s in sel):
        os.makedirs(os.path.join(HERE, 'extra'), exist_ok=True)
        golden_synthetic(ref, os.path.join(HERE, 'extra'),
                         only=[s for s in (sel or []) if s.startswith('xtr_')] or None, cases=EXTRA_CASES)
