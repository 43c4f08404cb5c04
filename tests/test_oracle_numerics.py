"""Every function of oracle/numerics_np.py against explicit Python loops over (k, p, q, i) written from
the mathematical definitions (the reference's own per-kernel tests are
tests/test.py:877-1217).  Small sizes, P = 1, 2, 3: the vectorised oracle and the loops must agree to
rounding.  CPU only; this pins the building blocks the trajectory goldens exercise as a whole.
"""
import numpy as np
import pytest

from oracle import numerics_np as nm

SHAPES = [(3, 1, 7, 1), (4, 2, 6, 2), (5, 3, 5, 3)]      # K, P, M, A


def problem(K, P, M, A, seed=0):
    rng = np.random.default_rng(seed)
    vi_mu = rng.normal(size=(K, P, M))
    logits = rng.normal(size=(M, K))
    vi_delta = np.exp(logits) / np.exp(logits).sum(axis=1, keepdims=True)
    hyper = rng.dirichlet(np.ones(K), size=A)
    ann = rng.integers(0, A, size=M)
    covs = []
    for k in range(K):
        a = rng.normal(size=(P, P))
        covs.append(a @ a.T + 0.5 * np.eye(P))
    prec = np.stack([np.linalg.inv(c) for c in covs])[..., None]          # [K,P,P,1]
    sld = rng.uniform(0.5, 3.0, size=(P, M))
    lam = np.zeros((K, P, P, M))
    for k in range(K):
        for i in range(M):
            lam[k, :, :, i] = prec[k, :, :, 0] + np.diag(sld[:, i])
    return dict(rng=rng, vi_mu=vi_mu, vi_delta=vi_delta, hyper=hyper, ann=ann, covs=covs, prec=prec,
                sld=sld, lam=lam)


@pytest.mark.parametrize('K,P,M,A', SHAPES)
def test_elementwise_and_moments(K, P, M, A):
    d = problem(K, P, M, A)
    rng = d['rng']
    old, new = rng.normal(size=(K, P, M)), rng.normal(size=(K, P, M))
    assert np.allclose(nm.sum_betas(old, new, 0.3), 0.3 * new + 0.7 * old)
    w, x, y, z = (rng.uniform(0.5, 2, size=(P, M)) for _ in range(4))
    assert np.allclose(nm.linked_ests(w, x, y, z), w / x - y * z)
    pm = np.zeros((P, M))
    for p in range(P):
        for i in range(M):
            pm[p, i] = sum(d['vi_delta'][i, k] * d['vi_mu'][k, p, i] for k in range(K))
    assert np.allclose(nm.posterior_mean(d['vi_mu'], d['vi_delta']), pm)
    sigma = nm.vi_sigma_inv(d['lam'])
    diag = np.einsum('kppi->kpi', sigma)
    pv = np.zeros((P, M))
    for p in range(P):
        for i in range(M):
            pv[p, i] = sum(d['vi_delta'][i, k] * (sigma[k, p, p, i] + d['vi_mu'][k, p, i]**2)
                           for k in range(K)) - pm[p, i]**2
    assert np.allclose(nm.pmv(pm, d['vi_mu'], d['vi_delta'], diag), pv)
    assert np.all(pv > 0)


@pytest.mark.parametrize('K,P,M,A', SHAPES)
def test_inverse_logdet_and_natural_products(K, P, M, A):
    d = problem(K, P, M, A)
    sigma = nm.vi_sigma_inv(d['lam'])
    logdet = nm.vi_sigma_log_det(sigma)
    for k in range(K):
        for i in range(M):
            assert np.allclose(sigma[k, :, :, i] @ d['lam'][k, :, :, i], np.eye(P), atol=1e-12)
            assert np.isclose(logdet[k, i], np.log(np.linalg.det(sigma[k, :, :, i])))
    nat_sigma = -0.5 * d['lam']
    eta = np.zeros((K, P, M))
    back = np.zeros((K, P, M))
    for k in range(K):
        for i in range(M):
            eta[k, :, i] = d['lam'][k, :, :, i] @ d['vi_mu'][k, :, i]
    assert np.allclose(nm.nat_inner_product_m2(d['vi_mu'], nat_sigma), eta)
    for k in range(K):
        for i in range(M):
            back[k, :, i] = sigma[k, :, :, i] @ eta[k, :, i]
    assert np.allclose(nm.nat_inner_product(eta, sigma), back)
    assert np.allclose(back, d['vi_mu'])


@pytest.mark.parametrize('K,P,M,A', SHAPES)
def test_kl_terms_and_likelihood(K, P, M, A):
    d = problem(K, P, M, A)
    rng = d['rng']
    quad = 0.0
    for i in range(M):
        for k in range(K):
            m = d['vi_mu'][k, :, i]
            quad += 0.5 * d['vi_delta'][i, k] * (m @ d['prec'][k, :, :, 0] @ m)
    assert np.isclose(nm.inner_product_comp(d['vi_mu'], d['prec'], d['vi_delta']), quad)
    kl = 0.0
    for i in range(M):
        for k in range(K):
            kl += d['vi_delta'][i, k] * (np.log(d['vi_delta'][i, k]) - np.log(d['hyper'][d['ann'][i], k]))
    assert np.isclose(nm.delta_kl(d['vi_delta'], d['hyper'], d['ann']), kl)
    summary = rng.normal(size=(M, K))
    assert np.isclose(nm.beta_kl(summary, d['vi_delta']), 0.5 * sum(
        summary[i, k] * d['vi_delta'][i, k] for i in range(M) for k in range(K)))
    sums = np.zeros((A, K))
    for i in range(M):
        sums[d['ann'][i]] += d['vi_delta'][i]
    assert np.allclose(nm.sum_annotations(d['vi_delta'], d['ann'], A), sums)
    # expected log-likelihood (numerics.py:31-46)
    pm, pv, z, linked, adj = (rng.normal(size=(P, M)) for _ in range(5))
    pv = np.abs(pv)
    chi, ranks, tau = rng.uniform(1, 5, size=P), rng.uniform(3, 9, size=P), rng.uniform(0.5, 2, size=P)
    want = 0.0
    for p in range(P):
        acc = sum(-0.5 * (d['sld'][p, i] * pv[p, i] + linked[p, i] * z[p, i]) + pm[p, i] * adj[p, i]
                  for i in range(M)) - 0.5 * chi[p]
        want += acc / tau[p] - 0.5 * ranks[p] * np.log(tau[p])
    assert np.isclose(nm.likelihood(pm, pv, z, d['sld'], linked, adj, chi, ranks, tau), want)


@pytest.mark.parametrize('K,P,M,A', SHAPES)
def test_categorical_maps(K, P, M, A):
    d = problem(K, P, M, A)
    rng = d['rng']
    log_det = np.array([np.linalg.slogdet(c)[1] for c in d['covs']])
    grad = np.zeros((M, max(K - 1, 0)))
    for i in range(M):
        a = d['ann'][i]
        last = np.log(d['hyper'][a, K - 1]) - 0.5 * log_det[K - 1]
        for k in range(K - 1):
            grad[i, k] = np.log(d['hyper'][a, k]) - 0.5 * log_det[k] - last
    assert np.allclose(nm.vi_delta_grad(d['hyper'], log_det, d['ann']), grad)
    nat = nm.map_to_nat_cat_2D(d['vi_delta'])
    for i in range(M):
        for k in range(K - 1):
            assert np.isclose(nat[i, k], np.log(d['vi_delta'][i, k] / d['vi_delta'][i, K - 1]))
    assert np.allclose(nm.invert_nat_cat_2D(nat), d['vi_delta'])
    # far-apart logits: floored at 1e-100, floored entries NOT renormalised (numerics.py:188-194)
    big = np.array([[800.0, -800.0], [0.0, -1e4]])
    out = nm.invert_nat_cat_2D(big)
    assert np.allclose(out[0], [1.0, 1e-100, 1e-100]) and np.allclose(out[1], [0.5, 1e-100, 0.5])
    # the update's softmax: logits 0.5[(c_k + mu.eta_k) - (c_K + mu.eta_K)] + g_k
    new_mu, nat_mu = rng.normal(size=(K, P, M)), rng.normal(size=(K, P, M))
    const, g = rng.normal(size=(M, K)), rng.normal(size=(M, K - 1))
    want = np.zeros((M, K))
    for i in range(M):
        t = np.array([const[i, k] + new_mu[k, :, i] @ nat_mu[k, :, i] for k in range(K)])
        logit = np.append(0.5 * (t[:-1] - t[-1]) + g[i], 0.0)
        e = np.exp(logit - logit.max())
        want[i] = np.maximum(e / e.sum(), 1e-100)
    assert np.allclose(nm.invert_nat_vi_delta(new_mu, nat_mu, const, g), want)
