"""`vilma sim`: simulate GWAS summary data from the mixture-of-Gaussians model (drop-in for
vilma.sim).

Same flags, inputs, RNG consumption order and output files as /root/reference/src/vilma/sim.py
(:11-70 parser, :161-272 driver).  The two LD products per cohort -- the mean S R S^-1 beta and the
noise S R^(1/2) eps (:140-158) -- run on the GPU through the same block-diagonal operator as the fit
(`BlockDiagonalMatrix.dot` -> vb_ld_dot; `matrix_power(0.5)` rescales the eigenvalues of the factors).
The per-SNP component draw (:73-96: one `np.random.choice` per SNP in a Python loop) is vectorised
without changing a single draw: legacy `RandomState.choice(K, p=w)` is one uniform searched in the
cumulative weights, so M single draws equal one `random_sample(M)` searched per annotation.
"""
import logging
import pickle

import numpy as np
import pandas as pd

from . import load


def args(super_parser):
    parser = super_parser.add_parser(
        'sim',
        description='Simulate GWAS summary data from a mixture-of-gaussians model.',
        usage='vilma sim <options>',
    )
    parser.add_argument('--sumstats', required=True, type=str,
                        help='Comma-separated paths to summary statistics.')
    parser.add_argument('--covariance', required=True, type=str,
                        help='Path to .pkl file containing the covariance matrices for each '
                             'Gaussian component.')
    parser.add_argument('--weights', required=True, type=str,
                        help='Path to a .npy file containing a matrix of weights (num_annotations x '
                             'num_components), or a .npz file containing a fitted vilma model.')
    parser.add_argument('--gwas-n-scaling', required=False, type=str, default='1.',
                        help='Comma-separated list of values to use to scale the sample sizes for '
                             'each cohort.')
    parser.add_argument('--annotations', type=str, default='', help='Path to annotations file.')
    parser.add_argument('--output', required=True, type=str, help='Output path prefix.')
    parser.add_argument('--names', type=str, required=False,
                        help='Comma-separated names of the populations for the output. Defaults to '
                             '0, 1, ...')
    parser.add_argument('--ld-schema', required=True, type=str,
                        help='Comma-separated paths to LD panel schemas.')
    parser.add_argument('--seed', type=int, default=42, help='Seed for random number generation.')
    return parser


def sim_components(annotations, weights):
    """One-hot [num_snps, num_components]: row i has its one in column j with probability
    weights[annotation of i, j] (sim.py:73-96), with the reference's draws."""
    annotations = np.asarray(annotations)
    weights = np.asarray(weights, dtype=np.float64)
    num_snps = annotations.shape[0]
    which = np.argmax(annotations == 1, axis=1)
    uniforms = np.random.random_sample(num_snps)
    out = np.zeros((num_snps, weights.shape[1]))
    for a in range(weights.shape[0]):
        rows = np.where(which == a)[0]
        if len(rows) == 0:
            continue
        cdf = weights[a].cumsum()
        cdf /= cdf[-1]
        out[rows, cdf.searchsorted(uniforms[rows], side='right')] = 1
    return out


def sim_true_effects(annotations, weights, cov_mats):
    """[num_pops, num_snps] effects drawn from the zero-mean mixture (sim.py:99-137)."""
    cov_mats = np.asarray(cov_mats, dtype=np.float64)
    num_pops = cov_mats.shape[-1]
    one_hot_components = sim_components(annotations, weights)
    latent_effects = np.random.normal(loc=0, scale=1, size=(annotations.shape[0], num_pops))
    sqrt_covs = np.array([np.linalg.cholesky(mat) for mat in cov_mats])
    comp = np.argmax(one_hot_components, axis=1)
    # 'ip,ik,kqp->qi' with a one-hot k: the SNP's own Cholesky factor applied to its latent draw
    return np.einsum('ip,iqp->qi', latent_effects, sqrt_covs[comp])


def sim_gwas(true_beta, std_errs, ld_mat):
    """GWAS estimates S R S^-1 beta + S R^(1/2) eps (sim.py:140-158); both products on the GPU."""
    mean = std_errs * (ld_mat.dot(true_beta / std_errs))
    latent_noise = np.random.normal(loc=0, scale=1, size=true_beta.shape[0])
    true_noise = std_errs * (ld_mat.matrix_power(0.5)).dot(latent_noise)
    return mean + true_noise


def main(args):
    np.random.seed(args.seed)
    num_pops = len(args.sumstats.split(','))
    names = list(map(str, range(num_pops)))
    if args.names is not None:
        if args.names.count(',') != args.sumstats.count(','):
            raise ValueError('If --names are provided, one must be provided per sumstat file.')
        names = args.names.split(',')

    n_scales = np.ones(num_pops)
    n_scales[:] = np.array(list(map(float, args.gwas_n_scaling.split(','))))
    if not np.all(n_scales > 0):
        raise ValueError('--gwas-n-scaling must be all positive.')

    all_vars = [load.load_variant_list(f) for f in args.sumstats.split(',')]
    all_vars = pd.concat(all_vars, ignore_index=True).drop_duplicates(subset='ID', ignore_index=True)

    annotations, denylist = load.load_annotations(args.annotations, all_vars)
    annotations = np.array(annotations, dtype=np.float64)
    num_annotations = annotations.shape[1]
    # un-annotated variants get an annotation drawn in proportion to the annotated ones (:197-206)
    annotation_proportions = annotations.sum(axis=0).astype(np.float64)
    annotation_proportions /= annotation_proportions.sum()
    random_annots = np.random.choice(num_annotations, size=len(denylist), p=annotation_proportions,
                                     replace=True)
    annotations[denylist, :] = 0
    annotations[denylist, random_annots] = 1
    assert np.all(annotations.sum(axis=1) == 1)

    # only the standard errors of the sumstats are used; missing data gets SE 1e-100 (:208-237)
    std_errs = np.ones((num_pops, all_vars.shape[0])) * 1e-100
    ld_mats = []
    for idx, (sstats_file, n_scale, ld_schema_path) in enumerate(
            zip(args.sumstats.split(','), n_scales, args.ld_schema.split(','))):
        logging.info('Loading sumstats for population %s...', names[idx])
        these_sstats, missing = load.load_sumstats(sstats_file, all_vars)
        logging.info('Loading LD for population %s...', names[idx])
        ld_mat, this_missing_ld = load.load_ld_from_schema(
            ld_schema_path, variants=all_vars, denylist=missing, ldthresh=0.999999, mmap=True)
        ld_mats.append(ld_mat)
        keep_bool = np.ones(all_vars.shape[0], dtype=bool)
        keep_bool[missing] = False
        keep_bool[this_missing_ld] = False
        std_errs[idx, keep_bool] = (np.sqrt(1 / n_scale) * these_sstats.SE.loc[keep_bool])

    with open(args.covariance, 'rb') as pickle_file:
        cov_mats = np.array(pickle.load(pickle_file)[0])

    weights = np.load(args.weights)
    try:
        weights.files
        weights = weights['hyper_delta']
    except AttributeError:
        weights = np.array(weights)
    if weights.shape[0] != num_annotations:
        raise ValueError('The shape of the weights does not match the number of annotations.')
    if weights.shape[1] != len(cov_mats):
        raise ValueError('The shape of the weights does not match the number of covariance '
                         'matrices.')
    if not np.allclose(weights.sum(axis=1), 1.):
        raise ValueError('weights do not sum to 1 within each annotation.')

    true_effects = sim_true_effects(annotations, weights, cov_mats)
    sim_beta_hat = np.zeros((num_pops, all_vars.shape[0]))
    for p, (ld_mat, beta, std_vec) in enumerate(zip(ld_mats, true_effects, std_errs)):
        sim_beta_hat[p] = sim_gwas(beta, std_vec, ld_mat)

    for p in range(num_pops):
        logging.info('Saving results for cohort %s', names[p])
        to_save = all_vars.copy()
        to_save['SE'] = std_errs[p]
        to_save['BETA'] = sim_beta_hat[p]
        to_save['true_beta'] = true_effects[p]
        to_save.loc[to_save.SE < 1e-99, 'SE'] = np.nan
        to_save = to_save.dropna()
        to_save.to_csv(args.output + '.' + names[p] + '.simgwas.tsv', sep='\t', index=False)
