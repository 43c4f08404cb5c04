"""NumPy restatement of the reference's numeric kernels (ORACLE -- tests only).

Every function names the numba kernel in ``/root/reference/src/vilma/numerics.py``
whose result it must reproduce.  Shapes: K mixture components, P cohorts, M SNPs,
A annotations.  Host layouts are the reference's: ``vi_mu[K,P,M]``,
``vi_delta[M,K]``, ``hyper_delta[A,K]``, covariances ``[K,P,P,M]``.
"""
import numpy as np

EPSILON = 1e-100    # numerics.py:8


def sum_betas(old, new, step):
    """numerics.py:11-15  step*new + (1-step)*old"""
    return step * new + (1. - step) * old


def linked_ests(w, x, y, z):
    """numerics.py:24-28  w/x - y*z (the code is minus; its docstring says plus)"""
    return w / x - y * z


def likelihood(pm, pv, z, sld, linked, adj, chi, ranks, tau):
    """numerics.py:31-46  expected log-likelihood under q"""
    per_pop = (-0.5 * (sld * pv + linked * z) + pm * adj).sum(axis=1)
    per_pop = per_pop - 0.5 * chi
    return float((per_pop / tau - 0.5 * ranks * np.log(tau)).sum())


def posterior_mean(vi_mu, vi_delta):
    """numerics.py:49-57  sum_k delta[i,k] mu[k,p,i] -> [P,M]"""
    return np.einsum('kpi,ik->pi', vi_mu, vi_delta)


def pmv(mean, vi_mu, vi_delta, diag_sigma):
    """numerics.py:60-65  sum_k delta (S_pp + mu^2) - mean^2 -> [P,M]"""
    return np.einsum('kpi,ik->pi', diag_sigma + vi_mu**2, vi_delta) - mean**2


def nat_inner_product_m2(vi_mu, nat_sigma):
    """numerics.py:68-80  -2 * sum_q nat_sigma[k,p,q,i] mu[k,q,i]"""
    return -2. * np.einsum('kpqi,kqi->kpi', nat_sigma, vi_mu)


def nat_inner_product(nat_mu, vi_sigma):
    """numerics.py:83-95  sum_q S[k,p,q,i] eta[k,q,i]"""
    return np.einsum('kpqi,kqi->kpi', vi_sigma, nat_mu)


def inner_product_comp(vi_mu, mixture_prec, vi_delta):
    """numerics.py:98-115  0.5 sum_ik delta_ik mu_ki^T Prec_k mu_ki"""
    quad = np.einsum('kpi,kqi,kqp->ik', vi_mu, vi_mu, mixture_prec[..., 0])
    return 0.5 * float((quad * vi_delta).sum())


def sum_annotations(deltas, annotations, num_annotations):
    """numerics.py:118-129  per-annotation column sums of delta -> [A,K]"""
    out = np.zeros((num_annotations, deltas.shape[1]))
    for a in range(num_annotations):
        out[a] = deltas[annotations == a].sum(axis=0)
    return out


def delta_kl(vi_delta, hyper_delta, annotations):
    """numerics.py:132-141  sum_ik delta (log delta - log hyper[a_i,k])"""
    log_hyper = np.log(hyper_delta)
    return float((vi_delta * (np.log(vi_delta) - log_hyper[annotations])).sum())


def beta_kl(sigma_summary, vi_delta):
    """numerics.py:144-146"""
    return 0.5 * float((sigma_summary * vi_delta).sum())


def vi_delta_grad(hyper_delta, log_det, annotations):
    """numerics.py:149-164  natural gradient of the categorical q -> [M,K-1]"""
    t = np.log(hyper_delta) - 0.5 * log_det          # [A,K]
    t = t[:, :-1] - t[:, -1:]
    return t[annotations]


def map_to_nat_cat_2D(probs):
    """numerics.py:167-176  log(p_k / p_last)"""
    lp = np.log(probs)
    return lp[:, :-1] - lp[:, -1:]


def invert_nat_cat_2D(nat):
    """numerics.py:179-195  softmax with implicit last logit 0, floored at EPSILON
    (floored entries are NOT renormalised)."""
    mx = np.maximum(nat.max(axis=1), 0.)
    e = np.exp(nat - mx[:, None])
    last = np.exp(-mx)
    denom = last + e.sum(axis=1)
    out = np.empty((nat.shape[0], nat.shape[1] + 1))
    out[:, :-1] = e / denom[:, None]
    out[:, -1] = last / denom
    return np.maximum(out, EPSILON)


def invert_nat_vi_delta(new_mu, nat_mu, const_part, nat_vi_delta):
    """numerics.py:198-213  logits 0.5[(c_ik + mu.eta_k) - (c_iK + mu.eta_K)] + g_ik"""
    t = const_part + np.einsum('kpi,kpi->ik', new_mu, nat_mu)
    return invert_nat_cat_2D(0.5 * (t[:, :-1] - t[:, -1:]) + nat_vi_delta)


def vi_sigma_inv(mats):
    """numerics.py:216-254  batched inverse of [K,P,P,M] over the middle two axes"""
    P = mats.shape[1]
    if P == 1:
        return 1. / mats
    if P == 2:
        a, b, c, d = mats[:, 0, 0], mats[:, 0, 1], mats[:, 1, 0], mats[:, 1, 1]
        idet = 1. / (a * d - b * c)
        out = np.empty_like(mats)
        out[:, 0, 0] = d * idet
        out[:, 1, 1] = a * idet
        # the reference mirrors the [1,0] entry of its transposed view into [0,1]
        out[:, 1, 0] = -c * idet
        out[:, 0, 1] = out[:, 1, 0]
        return out
    inv = np.linalg.inv(np.transpose(mats, (3, 0, 1, 2)))
    return np.transpose(inv, (1, 2, 3, 0))


def vi_sigma_log_det(mats):
    """numerics.py:257-290  batched log-determinant -> [K,M]"""
    P = mats.shape[1]
    if P == 1:
        return np.log(mats[:, 0, 0])
    if P == 2:
        return np.log(mats[:, 0, 0] * mats[:, 1, 1] - mats[:, 0, 1] * mats[:, 1, 0])
    ld = np.linalg.slogdet(np.transpose(mats, (3, 0, 1, 2)))[1]
    return ld.T
