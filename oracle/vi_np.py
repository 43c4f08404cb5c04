"""NumPy restatement of the reference fitting loop (ORACLE -- tests only).

Follows ``/root/reference/src/vilma/variational_inference.py`` (``VIScheme`` :27-564
and ``MultiPopVI`` :567-889); maths as in SURVEY.md Appendix A.  The control flow
(line search, L schedule, accumulated-delta ELBO, tau-step gate, convergence rule)
is restated literally because parity is judged on the decision sequence, not only
on the final numbers.  Redundant objective evaluations of the reference are kept
(this is the checker and the CPU baseline, not the product).
"""
import logging

import numpy as np

from . import numerics_np as nm

L_MAX = 1e12            # variational_inference.py:18
REL_TOL = 1e-6          # :19
ABS_TOL = 1e-6          # :20
ELBO_TOL = 0.1          # :21
EM_TOL = 10             # :22
ELBO_MOMENTUM = 0.5     # :23
MAX_NUM_ITERS = 20      # :24


class OracleVI:
    """Restatement of ``MultiPopVI`` (host NumPy, float64)."""

    param_names = ['vi_mu', 'vi_delta', 'hyper_delta']

    def __init__(self, marginal_effects, std_errs, ld_mats, mixture_covs,
                 annotations, scaled=False, scale_se=False, gwas_N=None,
                 init_hg=None, num_its=None, checkpoint=False,
                 checkpoint_freq=5, output='vilma_output'):
        # --- VIScheme.__init__ :96-259 -----------------------------------
        if not np.all(np.isfinite(marginal_effects)):
            raise ValueError('non-finite GWAS effect size')
        if not np.all(np.isfinite(std_errs)):
            raise ValueError('non-finite GWAS standard error')
        mixture_covs = [np.asarray(c, dtype=float) for c in mixture_covs]
        P, M = marginal_effects.shape
        for c in mixture_covs:
            if c.shape != (P, P):
                raise ValueError('Mixture component has a covariance matrix of '
                                 'the wrong shape.')
        if not np.all(np.linalg.slogdet(np.array(mixture_covs))[0] == 1):
            raise ValueError('Mixture component has a non-positive definite '
                             'covariance matrix.')
        self.num_pops, self.num_loci, self.num_mix = P, M, len(mixture_covs)
        self.scaled, self.scale_se = scaled, scale_se
        self.error_scaling = np.ones(P)
        self.checkpoint, self.checkpoint_freq = checkpoint, checkpoint_freq
        self.checkpoint_path = '%s-checkpoint' % output
        self.num_its = num_its
        self.ld_mats = ld_mats
        if len(ld_mats) != P:
            raise ValueError('Fewer LD matrices than populations.')
        self.ld_diags = np.stack([ld.diag() for ld in ld_mats])
        if not np.allclose(annotations.sum(axis=1), 1):
            raise ValueError('annotations must be one-hot')
        self.num_annotations = annotations.shape[1]
        eff = np.array(marginal_effects, dtype=float)
        se_in = np.array(std_errs, dtype=float)
        if scaled:                                              # :205-214
            eff = eff / (se_in + nm.EPSILON)
            self.std_errs = np.ones_like(se_in)
            self.scalings = se_in + nm.EPSILON
        else:
            self.std_errs = se_in
            self.scalings = np.ones_like(se_in)
        self.marginal_effects = eff
        self.scaled_ld_diags = self.std_errs**-2 * self.ld_diags
        self.annotations = np.where(annotations)[1].astype(np.int64)
        self.annotation_counts = annotations.sum(axis=0)
        self.adj_marginal_effects = np.zeros((P, M))
        self.chi_stat = np.zeros(P)
        self.ld_ranks = np.zeros(P)
        self.inverse_betas = np.zeros((P, M))
        for p in range(P):                                      # :236-252
            z = eff[p] / self.std_errs[p]
            mle = ld_mats[p].inverse.dot(z)
            self.chi_stat[p] = z.dot(mle)
            adj = ld_mats[p].dot(mle) / self.std_errs[p]
            self.adj_marginal_effects[p] = adj
            self.ld_ranks[p] = ld_mats[p].get_rank()
            prior = 2 * gwas_N[p] * init_hg[p] / (self.std_errs[p]**-2).sum()
            ridge = ld_mats[p].ridge_inverse_dot(adj * self.std_errs[p],
                                                 self.std_errs[p]**2 / prior)
            self.inverse_betas[p] = ridge * self.std_errs[p]
        if not np.allclose(self.adj_marginal_effects[np.isclose(self.ld_diags, 0)], 0):
            raise ValueError('Some SNPs that are missing in the LD matrix are '
                             'not being treated as missing.')
        # --- MultiPopVI.__init__ :599-630 ----------------------------------
        covs4 = np.array(mixture_covs)[:, :, :, None]
        self.mixture_prec = nm.vi_sigma_inv(covs4)              # [K,P,P,1]
        self.log_det = np.copy(nm.vi_sigma_log_det(covs4)[:, 0])
        self._set_vi_sigma()
        self.nat_grad_vi_delta = None
        self.counters = {'matvec': 0, 'trials': 0, 'beta_calls': 0}

    # ------------------------------------------------------------------
    def _set_vi_sigma(self):
        """:712-733  S_ki = (Prec_k + diag(sld_i / tau))^-1 and its summaries"""
        K, P, M = self.num_mix, self.num_pops, self.num_loci
        lam = np.zeros((K, P, P, M))
        idx = np.arange(P)
        lam[:, idx, idx, :] = self.scaled_ld_diags / self.error_scaling[:, None]
        lam += self.mixture_prec
        self.vi_sigma = nm.vi_sigma_inv(lam)
        self.nat_sigma = -0.5 * lam
        self.vi_sigma_log_det = nm.vi_sigma_log_det(self.vi_sigma)
        self.vi_sigma_matches = np.einsum('kpq,kqpi->ik', self.mixture_prec[..., 0],
                                          self.vi_sigma)
        self.sigma_summary = (self.log_det - self.vi_sigma_log_det.T
                              + self.vi_sigma_matches)

    def _matvec(self, p, x):
        self.counters['matvec'] += 1
        return self.ld_mats[p].dot(x)

    # ---- moments ------------------------------------------------------
    def _posterior_mean(self, vi_mu, vi_delta, hyper_delta=None):
        return nm.posterior_mean(vi_mu, vi_delta)               # :753-755

    def _posterior_marginal_variance(self, mean, vi_mu, vi_delta, hyper_delta=None):
        diag = np.einsum('kppi->kpi', self.vi_sigma)            # :757-760
        return nm.pmv(mean, vi_mu, vi_delta, diag)

    def real_posterior_mean(self, vi_mu, vi_delta, hyper_delta=None):
        return nm.posterior_mean(vi_mu, vi_delta) * self.scalings      # :740-743

    def real_posterior_variance(self, vi_mu, vi_delta, hyper_delta=None):
        mean = self._posterior_mean(vi_mu, vi_delta)
        return (self._posterior_marginal_variance(mean, vi_mu, vi_delta)
                * self.scalings**2)                             # :745-751

    # ---- objective ----------------------------------------------------
    def _log_likelihood(self, params):
        """:452-470"""
        pm = self._posterior_mean(*params)
        pv = self._posterior_marginal_variance(pm, *params)
        z = pm / self.std_errs
        linked = np.stack([self._matvec(p, z[p]) for p in range(self.num_pops)])
        return nm.likelihood(pm, pv, z, self.scaled_ld_diags, linked,
                             self.adj_marginal_effects, self.chi_stat,
                             self.ld_ranks, self.error_scaling)

    def _beta_KL(self, vi_mu, vi_delta, hyper_delta):
        """:873-885"""
        return (nm.delta_kl(vi_delta, hyper_delta, self.annotations)
                + nm.inner_product_comp(vi_mu, self.mixture_prec, vi_delta)
                + nm.beta_kl(self.sigma_summary, vi_delta))

    def elbo(self, params):
        """:412-417 (annotation KL is identically 0, :887-889)"""
        return self._log_likelihood(params) - self._beta_KL(*params)

    _beta_objective = elbo                                      # :488-490

    # ---- parameter maps -------------------------------------------------
    def _nat_to_not_vi_delta(self, params):
        """:632-641"""
        vi_mu, vi_delta, hyper_delta = params
        nat_mu = nm.nat_inner_product_m2(vi_mu, self.nat_sigma)
        vi_delta = nm.invert_nat_vi_delta(vi_mu, nat_mu, self.vi_sigma_log_det.T,
                                          self.nat_grad_vi_delta)
        return vi_mu, vi_delta, hyper_delta

    def _set_state(self, params):
        """:702-710"""
        self._set_vi_sigma()
        self.nat_grad_vi_delta = nm.vi_delta_grad(params[2], self.log_det,
                                                  self.annotations)

    def _initialize(self):
        """:643-700  seeded, jittered start from the ridge estimate"""
        real_mu = self.inverse_betas
        missing = np.isclose(self.ld_diags, 0)
        fake_mu = np.random.normal(loc=np.copy(real_mu), scale=1e-3 * self.std_errs,
                                   size=real_mu.shape)
        fake_mu[missing] = np.nan
        with np.errstate(all='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', category=RuntimeWarning)
                fill = np.tile(np.nanmean(fake_mu, axis=0), [fake_mu.shape[0], 1])
        fake_mu[missing] = fill[missing]
        fake_mu[np.isnan(fake_mu)] = 0.
        probs = np.einsum('pi,oi,kpo->ik', 1.6 * fake_mu, 1.6 * fake_mu,
                          self.mixture_prec[..., 0])
        probs += self.vi_sigma_matches
        probs -= self.log_det
        probs = np.exp(-0.5 * (probs - probs.min(axis=1, keepdims=True)))
        vi_delta = np.maximum(probs / probs.sum(axis=1, keepdims=True), nm.EPSILON)
        hyper = nm.sum_annotations(vi_delta, self.annotations, self.num_annotations)
        hyper += 1.
        hyper /= hyper.sum(axis=1, keepdims=True)
        hyper = np.maximum(hyper, nm.EPSILON)
        self.nat_grad_vi_delta = nm.vi_delta_grad(hyper, self.log_det, self.annotations)
        avg = np.einsum('kpqi,ik->ipq', self.vi_sigma, vi_delta)
        inv_avg = np.linalg.inv(avg)
        temp_nat_mu = np.einsum('pi,iqp->qi', fake_mu, inv_avg)
        vi_mu = np.einsum('kqpi,pi->kqi', self.vi_sigma, temp_nat_mu)
        _, vi_delta, _ = self._nat_to_not_vi_delta((vi_mu, vi_delta, hyper))
        return vi_mu, vi_delta, hyper

    # ---- updates --------------------------------------------------------
    def _nat_grad_beta(self, vi_mu, vi_delta, hyper_delta):
        """:804-823  g[p,i] (identical for every k) -- returned un-broadcast"""
        pm = self._posterior_mean(vi_mu, vi_delta)
        z = pm / self.std_errs
        linked = np.stack([self._matvec(p, z[p]) for p in range(self.num_pops)])
        linked = nm.linked_ests(linked, self.std_errs, pm, self.scaled_ld_diags)
        return (self.adj_marginal_effects - linked) / self.error_scaling[:, None]

    def _update_beta(self, vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
        """:762-802  natural-gradient step with backtracking on 1/L"""
        self.counters['beta_calls'] += 1
        if orig_obj is None:
            orig_obj = self._beta_objective((vi_mu, vi_delta, hyper_delta))
        old_nat_mu = nm.nat_inner_product_m2(vi_mu, self.nat_sigma)
        const_part = self.vi_sigma_log_det.T
        if self.nat_grad_vi_delta is None:
            raise RuntimeError('nat_grad_vi_delta must always be set prior to '
                               'running _update_beta')
        grad = self._nat_grad_beta(vi_mu, vi_delta, hyper_delta)[None]
        while True:
            step = 1. / L[idx]
            nat_mu = nm.sum_betas(old_nat_mu, grad, step)
            new_mu = nm.nat_inner_product(nat_mu, self.vi_sigma)
            new_delta = nm.invert_nat_vi_delta(new_mu, nat_mu, const_part,
                                               self.nat_grad_vi_delta)
            self.counters['trials'] += 1
            new_obj = self._beta_objective((new_mu, new_delta, hyper_delta))
            logging.info('...Old objective = %f, new objective = %f', orig_obj, new_obj)
            if new_obj >= orig_obj - REL_TOL * np.abs(orig_obj) - ABS_TOL:
                if L[idx] > L_MAX and not np.isclose(orig_obj, new_obj):
                    raise RuntimeError('Encountered a numerical error.')
                break
            if L[idx] > L_MAX:
                if not np.isclose(orig_obj, new_obj):
                    raise RuntimeError('Encountered a numerical error.')
                return (vi_mu, vi_delta, hyper_delta), L, orig_obj, orig_obj
            L[idx] *= lsr
        return (new_mu, new_delta, hyper_delta), L, orig_obj, new_obj

    def _update_hyper_delta(self, vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
        """:825-860  closed-form mixture-weight update, then refresh delta"""
        if orig_obj is None:
            orig_obj = self.elbo((vi_mu, vi_delta, hyper_delta))
        new_hyper = nm.sum_annotations(vi_delta, self.annotations, self.num_annotations)
        new_hyper = np.maximum(
            new_hyper / (self.annotation_counts.reshape((-1, 1)) + nm.EPSILON),
            nm.EPSILON)
        new_hyper /= new_hyper.sum(axis=1, keepdims=True)
        self.nat_grad_vi_delta = nm.vi_delta_grad(new_hyper, self.log_det,
                                                  self.annotations)
        _, new_delta, _ = self._nat_to_not_vi_delta((vi_mu, vi_delta, new_hyper))
        new_obj = self.elbo((vi_mu, new_delta, new_hyper))
        return (vi_mu, new_delta, new_hyper), L, orig_obj, new_obj

    def _update_annotation(self, vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
        """:862-866 no-op"""
        return (vi_mu, vi_delta, hyper_delta), L, 0., 0.

    def _update_error_scaling(self, params):
        """:472-486 and :735-738"""
        pm = self._posterior_mean(*params)
        pv = self._posterior_marginal_variance(pm, *params)
        new = np.zeros_like(self.error_scaling)
        for p in range(self.num_pops):
            z = pm[p] / self.std_errs[p]
            new[p] = (self.chi_stat[p] - 2 * pm[p].dot(self.adj_marginal_effects[p])
                      + z.dot(self._matvec(p, z))
                      + (self.ld_diags[p] * pv[p] * self.std_errs[p]**-2).sum()
                      ) / self.ld_ranks[p]
        self.error_scaling = new
        self._set_vi_sigma()

    # ---- loop -----------------------------------------------------------
    def _nat_grad_step(self, params, L, line_search_rate, running_elbo_delta=None):
        """:419-450"""
        updates = [self._update_beta, self._update_hyper_delta, self._update_annotation]
        conv_tol = float('inf') if running_elbo_delta is None else 0.1 * running_elbo_delta
        delta = 0
        for idx, update in enumerate(updates):
            orig_obj = None
            for _ in range(MAX_NUM_ITERS):
                L[idx] = max([1., L[idx] / 1.25])
                params, L, orig_obj, new_obj = update(*params, orig_obj, L, idx,
                                                      line_search_rate)
                delta += new_obj - orig_obj
                with np.errstate(invalid='ignore'):
                    small = np.isclose(new_obj - orig_obj, 0, atol=conv_tol, rtol=0)
                if small or L[idx] == 1 or L[idx] > L_MAX:
                    break
                orig_obj = new_obj
        if self.scale_se and delta < EM_TOL:
            orig_obj = self.elbo(params)
            self._update_error_scaling(params)
            params = self._nat_to_not_vi_delta(params)
            new_obj = self.elbo(params)
            delta += new_obj - orig_obj
        return params, L, delta

    def _optimize_step(self, params, L, curr_elbo, line_search_rate=1.25,
                       running_elbo_delta=None):
        """:396-410  ELBO is accumulated from deltas, never recomputed"""
        new_params, L_new, change = self._nat_grad_step(params, L, line_search_rate,
                                                        running_elbo_delta)
        elbo = curr_elbo + change
        if running_elbo_delta is None:
            running_elbo_delta = change
        running_elbo_delta *= ELBO_MOMENTUM
        running_elbo_delta += (1 - ELBO_MOMENTUM) * np.maximum(change, 0)
        return new_params, L_new, elbo, running_elbo_delta

    def create_dump_dict(self, params):
        """:333-338"""
        d = dict(zip(self.param_names, params))
        d['error_scaling'] = self.error_scaling
        d['scalings'] = self.scalings
        return d

    def optimize(self, loaded_checkpoint=None, trajectory=None):
        """:340-394.  `trajectory` (a dict of lists) records per-iteration values."""
        if loaded_checkpoint is None:
            params = self._initialize()
        else:
            params = [np.array(loaded_checkpoint[n]) for n in self.param_names]
            try:
                self.error_scaling = np.array(loaded_checkpoint['error_scaling'])
            except KeyError:
                logging.warning('Did not find "error_scaling" in the loaded checkpoint.')
            self._set_state(params)
        converged = False
        elbo = self.elbo(params)
        running = None
        it = 0
        L = np.ones(5)
        post_mean = self.real_posterior_mean(*params)
        while it < self.num_its and not converged:
            if self.checkpoint and it % self.checkpoint_freq == 0:
                np.savez('{}.{}'.format(self.checkpoint_path, it),
                         **self.create_dump_dict(params))
            t0 = self.counters['trials']
            new_params, L, elbo, running = self._optimize_step(
                params, L=L, curr_elbo=elbo, line_search_rate=2.,
                running_elbo_delta=running)
            new_pm = self.real_posterior_mean(*new_params)
            converged = np.allclose(new_pm, post_mean, atol=ABS_TOL, rtol=REL_TOL)
            converged = converged or np.isclose(running, 0, atol=ELBO_TOL, rtol=0)
            if it < 10 and loaded_checkpoint is None:
                converged = False
            if trajectory is not None:
                trajectory.setdefault('elbo_out', []).append(float(elbo))
                trajectory.setdefault('L0', []).append(float(L[0]))
                trajectory.setdefault('trials', []).append(self.counters['trials'] - t0)
                trajectory.setdefault('tau', []).append(np.array(self.error_scaling))
                trajectory.setdefault('running', []).append(float(running))
            post_mean = new_pm
            it += 1
            params = tuple(new_params)
        if it == self.num_its:
            logging.warning('Failed to converge')
        return params
