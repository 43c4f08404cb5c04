"""Set-up on the GPU: `vb_setup_dense` through
BlockDiagonalMatrix.device_setup against the reference's host algebra -- eigendecomposition
(matrix_structures.py:15-28), pseudo-inverse product (:159-196), Woodbury ridge solve (:349-387) --
restated by the oracle (oracle/ld_np.py), and the constructor values of a golden fixture."""
import numpy as np
import pytest

from _fixtures import load_case, vi_kwargs

pytestmark = pytest.mark.gpu


def ar1_corr(n, n_ref, rng, rho=0.9):
    e = rng.standard_normal((n_ref, n))
    g = np.empty_like(e)
    g[:, 0] = e[:, 0]
    for j in range(1, n):
        g[:, j] = rho * g[:, j - 1] + np.sqrt(1 - rho * rho) * e[:, j]
    g += np.sqrt(0.1) * rng.standard_normal((n_ref, n))
    g -= g.mean(axis=0)
    g /= np.sqrt((g * g).sum(axis=0))
    r = g.T @ g
    r = 0.5 * (r + r.T)
    np.fill_diagonal(r, 1.0)
    return r


def oracle_setup(mats, perm, missing, z, reg):
    from oracle.ld_np import BlockDiagonalLD, LowRankBlock
    ld = BlockDiagonalLD([LowRankBlock(X=m, t=1.0) for m in mats], perm=perm, missing=missing)
    mle = ld.inverse.dot(z)
    rmle = ld.dot(mle)
    return mle, rmle, ld.ridge_inverse_dot(rmle, reg), float(z.dot(mle)), ld.get_rank()


def test_device_setup_matches_host_algebra():
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    rng = np.random.default_rng(4)
    sizes = [1, 7, 33, 64, 65, 130, 257, 700, 96]
    mats = [ar1_corr(n, 2 * n + 3, rng) if n > 1 else np.ones((1, 1)) for n in sizes]
    tot = sum(sizes)
    M = tot + 6
    order = rng.permutation(M)
    perm, missing = order[:tot], np.sort(order[tot:])
    full_perm = np.concatenate([perm, missing])
    z = rng.standard_normal(M)
    z[missing] = 0.0
    reg = rng.uniform(0.5, 30.0, size=M)
    ld = BlockDiagonalMatrix([LowRankMatrix(X=m, t=1.0, lazy=True) for m in mats], perm=full_perm,
                             missing=missing)
    res = ld.device_setup(z, reg)
    assert res['gpu_blocks'] == len(sizes) and res['host_blocks'] == 0
    assert not any(m.factorized for m in ld.matrices)            # no eigendecomposition ran
    mle, rmle, ridge, chi, rank = oracle_setup(mats, full_perm, missing, z, reg)
    scale = lambda a: 1e-9 * np.abs(a).max()
    assert np.allclose(res['mle'], mle, rtol=1e-8, atol=scale(mle))
    assert np.allclose(res['rmle'], rmle, rtol=1e-9, atol=scale(rmle))
    assert np.allclose(res['ridge'], ridge, rtol=1e-9, atol=scale(ridge))
    assert np.isclose(res['chi'], chi, rtol=1e-9)
    assert res['rank'] == rank == tot
    assert np.all(res['mle'][missing] == 0) and np.all(res['ridge'][missing] == 0)
    # the operator uploaded afterwards is the matrix itself (still no eigh), and it is the same operator
    x = rng.standard_normal(M)
    y = ld.dot(x)
    assert not any(m.factorized for m in ld.matrices)
    ref = np.zeros(M)
    off = 0
    for n, m in zip(sizes, mats):
        idx = perm[off:off + n]
        ref[idx] = m @ x[idx]
        off += n
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    assert np.allclose(ld.diag()[perm], 1.0) and np.all(ld.diag()[missing] == 0)
    # deterministic
    again = ld.device_setup(z, reg)
    assert np.array_equal(again['ridge'], res['ridge']) and again['chi'] == res['chi']
    ld.release_device()


def test_device_setup_declines_rank_deficient_and_indefinite_blocks():
    """Blocks where the reference WOULD drop eigenpairs (rank-deficient sample LD, an indefinite matrix,
    a near-singular one) are sent to the exact host path; the well-conditioned ones stay on the GPU."""
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    rng = np.random.default_rng(9)
    good = ar1_corr(120, 300, rng)
    deficient = ar1_corr(90, 40, rng)                     # rank 39
    indef = ar1_corr(50, 120, rng)
    indef[0, 1] = indef[1, 0] = 1.2                       # a negative eigenvalue
    w, v = np.linalg.eigh(ar1_corr(60, 150, rng))
    w[0] = 1e-11 * w[-1]                                  # kept by the reference (> 1e-12 max) but too close to call
    near = (v * w) @ v.T
    near = 0.5 * (near + near.T)
    mats = [good, deficient, indef, near, good.copy()]
    sizes = [m.shape[0] for m in mats]
    M = sum(sizes)
    z = rng.standard_normal(M)
    reg = rng.uniform(1.0, 5.0, size=M)
    ld = BlockDiagonalMatrix([LowRankMatrix(X=m, t=1.0, lazy=True) for m in mats])
    res = ld.device_setup(z, reg)
    assert res['gpu_blocks'] == 2 and res['host_blocks'] == 3
    assert [m.full_rank_certified for m in ld.matrices] == [True, False, False, False, True]
    assert [m.factorized for m in ld.matrices] == [False, True, True, True, False]
    mle, rmle, ridge, chi, rank = oracle_setup(mats, np.arange(M), np.array([], dtype=np.int64), z, reg)
    assert res['rank'] == rank
    assert np.allclose(res['rmle'], rmle, rtol=1e-8, atol=1e-9 * np.abs(rmle).max())
    assert np.allclose(res['ridge'], ridge, rtol=1e-8, atol=1e-9 * np.abs(ridge).max())
    lo = sizes[0]
    hi = lo + sizes[1] + sizes[2]
    # (the near-singular block's pseudo-inverse amplifies rounding by 1e11: compare the others)
    keep = np.r_[0:lo + sizes[1] + sizes[2], M - sizes[4]:M]
    assert np.allclose(res['mle'][keep], mle[keep], rtol=1e-7, atol=1e-8 * np.abs(mle[keep]).max())
    assert hi > lo
    ld.release_device()


def test_constructor_through_gpu_setup_matches_reference_fixture():
    """MultiPopVI built on lazily loaded dense blocks (the `--ldthresh 1` load path) reproduces the
    reference constructor's adj_marginal_effects / chi_stat / ld_ranks / inverse_betas and the fit."""
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    fx = load_case('syn_p1_dense')
    blocks = []
    for b in range(int(fx['ld0_nblocks'])):
        u, s = fx['ld0_u%d' % b], fx['ld0_s%d' % b]
        x = (u * s) @ u.T
        blocks.append(LowRankMatrix(X=0.5 * (x + x.T), t=1.0, lazy=True))
    ld = BlockDiagonalMatrix(blocks, perm=fx['ld0_perm'], missing=fx['ld0_missing'])
    vi = MultiPopVI(ld_mats=[ld], **vi_kwargs(fx))
    assert vi.setup_report == [(len(blocks), 0)]
    assert not any(m.factorized for m in blocks)
    assert np.allclose(vi.adj_marginal_effects, fx['pre_adj_marginal_effects'], rtol=1e-8,
                       atol=1e-10 * np.abs(fx['pre_adj_marginal_effects']).max())
    assert np.allclose(vi.chi_stat, fx['pre_chi_stat'], rtol=1e-9)
    assert np.array_equal(vi.ld_ranks, fx['pre_ld_ranks'])
    assert np.allclose(vi.inverse_betas, fx['pre_inverse_betas'], rtol=1e-7,
                       atol=1e-9 * np.abs(fx['pre_inverse_betas']).max())
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=0)
    assert np.allclose(vi.real_posterior_mean(*params), fx['final_post_mean'], rtol=1e-6, atol=1e-9)
    vi.close()
