"""`vilma fit`: argument parser and driver (drop-in for vilma.vi_options).

Same flags, defaults, input handling, mixture-grid construction, RNG consumption order and
output files as /root/reference/src/vilma/vi_options.py (:9-84 parser, :87-281 driver,
:284-337 grid); the fit itself runs on the GPU through ``MultiPopVI``.
"""
import itertools
import logging
import pickle

import numpy as np

from . import load


def args(super_parser):
    parser = super_parser.add_parser(
        'fit',
        description='Use variational inference to learn effect sizes and effect size '
                    'distribution from GWAS summary data.',
        usage='vilma fit <options>',
    )
    parser.add_argument('-K', '--components', default=12, type=int,
                        help='number of mixture components in prior')
    parser.add_argument('--num-its', default=1000, type=int,
                        help='Maximum number of optimization iterations.')
    parser.add_argument('--ld-schema', required=True, type=str,
                        help='Comma-separated paths to LD panel schemas.')
    parser.add_argument('--sumstats', required=True, type=str,
                        help='Comma-separated paths to summary statistics.')
    parser.add_argument('--stderrscale', default='1.0', type=str, required=False,
                        help='Comma separated list of values to multiply summary stat '
                             'stderrs by.')
    parser.add_argument('--annotations', type=str, default=None,
                        help='Path to annotation file.')
    parser.add_argument('--output', required=True, type=str, help='Output path prefix.')
    parser.add_argument('--names', type=str, required=False,
                        help='Comma-separated names of the populations for output. '
                             'Defaults to 0, 1,... ')
    parser.add_argument('--extract', required=True, type=str,
                        help='List of SNPs to include in analysis, with ID, A1, and A2 '
                             'columns.')
    parser.add_argument('--scaled', dest='scaled', action='store_true',
                        help='Place the prior on frequency-scaled effect sizes instead of '
                             'effect sizes in their natural scaling.')
    parser.add_argument('--ldthresh', required=False, default=1.0, type=float,
                        help='Threshold for singular value approximation of LD matrix. '
                             '--ldthresh x guarantees that SNPs with an r^2 of x or larger '
                             'will be linearly independent; 1 means no thresholding.')
    parser.add_argument('--seed', type=int, default=42,
                        help='Seed for random number generation.')
    parser.add_argument('--mmap', dest='mmap', action='store_true',
                        help='Accepted for compatibility; LD is kept in GPU memory.')
    parser.add_argument('--learn-scaling', dest='scale_se', action='store_true',
                        help='Whether or not to learn a scaling factor for the standard '
                             'errors.')
    parser.add_argument('--samplesizes', type=str, default='100e3',
                        help='Comma-separated GWAS sample sizes, used when initializing.')
    parser.add_argument('--init-hg', type=str, default='0.1',
                        help='Comma-separated heritabilities used when initializing.')
    parser.add_argument('--trait', dest='trait', action='store_true',
                        help='Treat sumstats files as different traits. Unimplemented.')
    parser.add_argument('--checkpoint-freq', type=int, default=-1,
                        help='Store the model once every this many iterations. Defaults '
                             'to no checkpointing.')
    parser.add_argument('--load-checkpoint', type=str, default='', nargs=2,
                        help='Resume from a saved checkpoint: the .npz file and the .pkl '
                             'file with the covariance matrices.',
                        metavar=('CHECKPOINT_FILE.npz', 'COVARIANCE_FILE.pkl'))
    return parser


def main(args):
    np.random.seed(args.seed)

    # (sic) the reference compares the comma count with 1, not 0 (vi_options.py:90-94)
    if (not args.trait
            and args.ld_schema.count(',') != 1
            and args.ld_schema.count(',') != args.sumstats.count(',')):
        raise ValueError('Either need to imput one ld_schema or provide a sumstats file '
                         'for each ld_schema.')

    num_pops = args.sumstats.count(',') + 1
    num_components = args.components
    names = list(map(str, range(num_pops)))
    if args.names is not None:
        if args.names.count(',') != args.sumstats.count(','):
            raise ValueError('If --names are provided, one must be provided per sumstat '
                             'file.')
        names = args.names.split(',')

    logging.info('Loading variants...')
    variants = load.load_variant_list(args.extract)
    logging.info('Loading annotations...')
    annotations, denylist = load.load_annotations(args.annotations, variants=variants)

    missing_annot = np.zeros(len(annotations), dtype=bool)
    missing_annot[denylist] = True
    missing_sumstats = np.zeros((len(annotations), num_pops), dtype=bool)
    missing_ld_info = np.zeros((len(annotations), num_pops), dtype=bool)

    combined_ld, combined_betas, combined_errors = [], [], []
    stderr_mult = np.zeros(len(args.sumstats.split(',')))
    stderr_mult[:] = list(map(float, args.stderrscale.split(',')))
    gwas_n = np.zeros_like(stderr_mult)
    gwas_n[:] = list(map(float, args.samplesizes.split(',')))
    init_hg = np.zeros_like(gwas_n)
    init_hg[:] = list(map(float, args.init_hg.split(',')))

    if args.trait:
        raise NotImplementedError('--trait has not been implemented yet.')
    for idx, (ld_schema_path, sumstats_path) in enumerate(
            zip(args.ld_schema.split(','), args.sumstats.split(','))):
        logging.info('Loading sumstats for population %d...', (idx + 1))
        sumstats, missing = load.load_sumstats(sumstats_path, variants=variants)
        missing_sumstats[missing, idx] = True
        missing.extend(denylist)
        combined_betas.append(np.array(sumstats.BETA).reshape((1, -1)))
        logging.info('Largest beta is... %f', np.max(np.abs(np.array(sumstats.BETA))))
        combined_errors.append(np.array(sumstats.SE).reshape((1, -1)) * stderr_mult[idx])
        logging.info('Loading LD for population %d...', (idx + 1))
        ld_mat, this_missing_ld = load.load_ld_from_schema(
            ld_schema_path, variants=variants, denylist=missing, ldthresh=args.ldthresh,
            mmap=args.mmap)
        combined_ld.append(ld_mat)
        missing_ld_info[this_missing_ld, idx] = True

    logging.info('Largest beta is... %f', np.max(np.abs(combined_betas)))
    betas = np.concatenate(combined_betas, axis=0)
    std_errs = np.concatenate(combined_errors, axis=0)

    if args.load_checkpoint:
        with open(args.load_checkpoint[1], 'rb') as pfile:
            cross_pop_covs = pickle.load(pfile)[0]
    else:
        logging.info('Building cross-population covariances...')
        mins, maxes = _grid_range(betas, std_errs, args.scaled)
        cross_pop_covs = _make_simple(num_pops, num_components, mins, maxes)
        from .dist import default_comm
        if default_comm().rank == 0:          # one writer when several ranks run the same command
            with open('%s.covariance.pkl' % args.output, 'wb') as ofile:
                pickle.dump([cross_pop_covs], ofile)

    logging.info('Fitting...')
    from .variational_inference import MultiPopVI
    elbo = MultiPopVI(
        marginal_effects=betas, std_errs=std_errs, ld_mats=combined_ld,
        mixture_covs=cross_pop_covs, annotations=annotations,
        checkpoint=(args.checkpoint_freq > 0), checkpoint_freq=args.checkpoint_freq,
        output=args.output, scaled=args.scaled, scale_se=args.scale_se, gwas_N=gwas_n,
        init_hg=init_hg, num_its=args.num_its,
    )
    checkpoint = None
    if args.load_checkpoint:
        checkpoint = np.load(args.load_checkpoint[0])
    params = elbo.optimize(checkpoint)

    rank0 = elbo._comm.rank == 0
    # np.savez(args.output, vi_mu, vi_delta, hyper_delta, error_scaling, scalings, vi_sigma) of the
    # reference (:263-265), with vi_sigma[K,P,P,M] streamed from the device slice by slice
    from .outputs import save_fit_npz
    save_fit_npz(args.output, elbo, params)

    for name, posterior in zip(names, elbo.real_posterior_mean(*params)):
        variants['posterior_' + name] = posterior
    for name, pmv in zip(names, elbo.real_posterior_variance(*params)):
        variants['posterior_variance_' + name] = pmv
    if args.annotations:
        variants['missing_annotation'] = missing_annot
    for idx, name in enumerate(names):
        variants['missing_sumstats_' + name] = missing_sumstats[:, idx]
        variants['missing_LD_' + name] = missing_ld_info[:, idx]
    if rank0:
        variants.to_csv(args.output + '.estimates.tsv', sep='\t', index=False)
    return elbo


def _grid_range(betas, std_errs, scaled):
    """Plausible smallest / largest true effect sizes per cohort (vi_options.py:196-226)."""
    num_pops = betas.shape[0]
    if scaled:
        maxes = np.nanmax((betas / std_errs)**2, axis=1)
        mins = np.zeros_like(maxes)
        for p in range(num_pops):
            keep = betas[p, :]**2 > 0
            mins[p] = np.nanpercentile((betas[p, keep] / std_errs[p, keep])**2, 2.5)
        return mins, maxes
    maxes = np.zeros(num_pops)
    mins = np.zeros_like(maxes)
    for p in range(num_pops):
        keep = ~np.isnan(betas[p])
        this_beta = np.abs(betas[p, keep])
        this_se = std_errs[p, keep]
        psi = 1. / len(this_beta)
        with np.errstate(over='ignore'):
            probs = 1. / (1. + ((1. - psi) / psi * np.sqrt(this_beta**2 / this_se**2)
                                * np.exp(-0.5 * this_beta**2 / this_se**2 + 0.5)))
        ebayes = np.maximum(this_beta**2 - this_se**2, 1e-10)
        raw_means = this_beta / (1. + this_se**2 / ebayes**2)
        maxes[p] = np.max(probs * raw_means)**2
        mins[p] = np.nanpercentile(betas[p, betas[p, :]**2 > 0]**2, 2.5)
    return mins, maxes


def _make_diag_vals(num_pops, num_components, mins, maxes):
    """Grid of variances across the populations (vi_options.py:284-298)."""
    diag_vals = [[m * 1e-6 for m in mins]]      # something that is basically zero
    for k in range(num_components + 1):
        diag_vals.append([mins[p] * np.exp(np.log(maxes[p] / mins[p]) / num_components * k)
                          for p in range(num_pops)])
    return diag_vals


def _make_simple(num_pops, num_components, mins, maxes):
    """Grid of covariance matrices (vi_options.py:301-337).  For num_pops > 1 it draws
    np.random.uniform three times per grid point, in this exact order."""
    diag_vals = _make_diag_vals(num_pops, num_components, mins, maxes)
    if num_pops == 1:
        return list(np.array(diag_vals).reshape((num_components + 2, num_pops, num_pops)))
    cross_pop_covs = []
    corr_vals = [-.99 + 1.98 * (k + 1) / num_components for k in range(num_components)]
    n_off = (num_pops * (num_pops - 1)) // 2

    def jitter(mat):
        for _ in range(3):
            scale = np.diag(np.sqrt(np.exp(np.random.uniform(-1, 1, num_pops))))
            cross_pop_covs.append(scale.dot(mat.dot(scale)))

    for idx, diag in enumerate(diag_vals):
        for off_diags in itertools.product(*[corr_vals] * n_off):
            mat = np.eye(num_pops)
            mat[np.triu_indices_from(mat, k=1)] = off_diags
            mat.T[np.triu_indices_from(mat, k=1)] = off_diags
            mat = mat * np.sqrt(diag)
            mat = mat.T * np.sqrt(diag)
            jitter(mat)
        if idx > 0:
            # population specific causals
            for population in range(num_pops):
                single_pop = np.copy(diag_vals[0])
                single_pop[population] = diag[population]
                jitter(np.diag(single_pop))
    return cross_pop_covs
