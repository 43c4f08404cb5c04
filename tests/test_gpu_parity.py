"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the goldens.

Tolerances (BASELINE.json north_star): posterior means / mixture weights rtol 1e-6 (with the
reference's own atol-style floor for near-zero entries), ELBO trajectory rtol 1e-8, identical
iteration count, line-search decisions (trial counts, L schedule).
"""
import numpy as np
import pytest

from _fixtures import VI_CASES, build_ld, load_case, vi_kwargs

pytestmark = pytest.mark.gpu


def make_product(fx, **extra):
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    kw = vi_kwargs(fx)
    kw.update(extra)
    return MultiPopVI(ld_mats=build_ld(fx, LowRankMatrix, BlockDiagonalMatrix), **kw)


@pytest.mark.parametrize('name', VI_CASES)
def test_ld_dot(name):
    """BlockDiagonalMatrix.dot on the GPU == oracle operator (perm, missing, low rank)."""
    from oracle.ld_np import BlockDiagonalLD, LowRankBlock
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    fx = load_case(name)
    ours = build_ld(fx, LowRankMatrix, BlockDiagonalMatrix)
    ref = build_ld(fx, LowRankBlock, BlockDiagonalLD)
    rng = np.random.default_rng(5)
    for a, b in zip(ours, ref):
        x = rng.standard_normal(a.shape[0])
        ya, yb = a.dot(x), b.dot(x)
        assert np.allclose(ya, yb, rtol=1e-12, atol=1e-12 * np.abs(yb).max())
        assert np.all(ya[a.missing] == 0)
        a.release_device()


@pytest.mark.parametrize('name', VI_CASES)
def test_precompute_and_state_eval(name):
    fx = load_case(name)
    vi = make_product(fx)
    assert np.allclose(vi.adj_marginal_effects, fx['pre_adj_marginal_effects'], rtol=1e-8,
                       atol=1e-10 * np.abs(fx['pre_adj_marginal_effects']).max())
    assert np.allclose(vi.chi_stat, fx['pre_chi_stat'], rtol=1e-9)
    assert np.array_equal(vi.ld_ranks, fx['pre_ld_ranks'])
    params = (fx['init_vi_mu'], fx['init_vi_delta'], fx['init_hyper_delta'])
    vi._set_state(params)
    assert np.isclose(vi.elbo(params), float(fx['init_elbo']), rtol=1e-11)
    assert np.isclose(vi._log_likelihood(params), float(fx['init_loglik']), rtol=1e-11)
    assert np.isclose(vi._beta_KL(*params), float(fx['init_beta_kl']), rtol=1e-11)
    assert np.allclose(vi.real_posterior_mean(*params), fx['init_post_mean'], rtol=1e-10, atol=1e-14)
    # pv = E[b^2] - E[b]^2 cancels catastrophically when |mean| >> sd (here up to E[b]^2/pv ~ 3e9):
    # the reference's own value is only accurate to eps * E[b^2] (it is 1.5e-7 relative away from a
    # long-double evaluation on this fixture), so that is the meaningful bound.
    pm_ref, pv_ref = fx['init_post_mean'], fx['init_post_var']
    pv = vi.real_posterior_variance(*params)
    assert np.all(np.abs(pv - pv_ref) <= 1e-10 * pv_ref + 64 * np.finfo(float).eps * (pv_ref + pm_ref**2))
    # seeded initialisation consumes the legacy RNG stream like the reference
    np.random.seed(int(fx['seed']))
    mu, delta, hyper = vi._initialize()
    assert np.allclose(mu, fx['init_vi_mu'], rtol=1e-7, atol=1e-12)
    assert np.allclose(delta, fx['init_vi_delta'], rtol=1e-7, atol=1e-300)
    assert np.allclose(hyper, fx['init_hyper_delta'], rtol=1e-9)
    assert np.allclose(vi.nat_grad_vi_delta, fx['init_nat_grad_vi_delta'], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('name', VI_CASES)
def test_initialize_on_device(name):
    """vb_init_delta_kernel / vb_init_mu_kernel (used when the host form's [K,P,P,M] arrays would not
    fit) give the reference's starting point, and the state they leave resident is that point."""
    fx = load_case(name)
    vi = make_product(fx)
    vi.init_on_device = True
    np.random.seed(int(fx['seed']))
    params = vi._initialize()
    assert np.allclose(params[0], fx['init_vi_mu'], rtol=1e-7, atol=1e-12)
    assert np.allclose(params[1], fx['init_vi_delta'], rtol=1e-7, atol=1e-300)
    assert np.allclose(params[2], fx['init_hyper_delta'], rtol=1e-9)
    assert np.allclose(vi.nat_grad_vi_delta, fx['init_nat_grad_vi_delta'], rtol=1e-9, atol=1e-12)
    assert vi._same(params)
    assert np.isclose(vi.elbo(params), float(fx['init_elbo']), rtol=1e-9)


@pytest.mark.parametrize('name', VI_CASES)
def test_trajectory(name):
    """Full optimize(): same decisions, same ELBO trajectory, same final parameters."""
    fx = load_case(name)
    vi = make_product(fx)
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    tr = vi.trajectory
    assert len(tr['elbo']) == len(fx['traj_elbo_out'])
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=0)
    assert np.allclose(vi.error_scaling, fx['final_error_scaling'], rtol=1e-8)
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=1e-9)
    assert np.allclose(params[1], fx['final_vi_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(params[2], fx['final_hyper_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(vi.real_posterior_mean(*params), fx['final_post_mean'], rtol=1e-6, atol=1e-9)
    assert np.allclose(vi.real_posterior_variance(*params), fx['final_post_var'], rtol=1e-6, atol=1e-12)
    assert np.allclose(vi.vi_sigma, fx['final_vi_sigma'], rtol=1e-8)
    # tracked ELBO == recomputed ELBO (reference tests/test.py:1574)
    assert np.isclose(vi.elbo(params), tr['elbo'][-1], rtol=1e-9)


@pytest.mark.parametrize('name', ['vischeme_linked_a2_s1_t1', 'syn_p1_dense', 'syn_p3'])
def test_trajectory_python_loop(name):
    """The Python control loop (used with INFO logging) takes the same decisions as the C++ one."""
    fx = load_case(name)
    vi = make_product(fx)
    vi.use_native_loop = False
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=0)
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=1e-9)
    assert np.allclose(vi.error_scaling, fx['final_error_scaling'], rtol=1e-8)


@pytest.mark.parametrize('warps', [1, 2, 8])
@pytest.mark.parametrize('name', ['vischeme_linked_a2_s1_t1', 'syn_p1_dense', 'syn_p1_scaled',
                                  'syn_p2_lowrank', 'syn_p3', 'syn_p5'])
def test_trajectory_tile_kernel(name, warps):
    """The K-split tile kernel (csrc/snp_tile_kernel.cuh), forced with W warps per 32-SNP tile, takes
    the reference's decisions too (by default it serves P >= 3 and K >= 32 only)."""
    from vilma_b200.engine import set_option
    fx = load_case(name)
    set_option('snp_tile', warps)
    try:
        vi = make_product(fx)
        np.random.seed(int(fx['seed']))
        params = vi.optimize(None)
    finally:
        set_option('snp_tile', -1)
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=0)
    assert np.allclose(vi.error_scaling, fx['final_error_scaling'], rtol=1e-8)
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=1e-9)
    assert np.allclose(params[1], fx['final_vi_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(params[2], fx['final_hyper_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(vi.real_posterior_mean(*params), fx['final_post_mean'], rtol=1e-6, atol=1e-9)
    assert np.allclose(vi.real_posterior_variance(*params), fx['final_post_var'], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize('option', ['snp3_park', 'snp_three_pass'])
@pytest.mark.parametrize('name', ['vischeme_linked_a2_s1_t1', 'syn_p1_dense', 'syn_p1_scaled'])
def test_trajectory_fallback_kernels(name, option):
    """The per-SNP kernels that the defaults no longer reach on these fixtures: the three-pass kernel
    parking its logits in the output buffers (snp3_park=0; used when K (P+1) KB exceeds 32 KB of
    shared memory) and the online single-pass kernel (snp_three_pass=0)."""
    from vilma_b200.engine import set_option
    fx = load_case(name)
    set_option(option, 0)
    try:
        vi = make_product(fx)
        np.random.seed(int(fx['seed']))
        params = vi.optimize(None)
    finally:
        set_option(option, 1)
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=0)
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=1e-9)
    assert np.allclose(params[1], fx['final_vi_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(params[2], fx['final_hyper_delta'], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize('name', [n for n in VI_CASES if 'resume_ckpt_vi_mu' in load_case(n)])
def test_resume(name):
    fx = load_case(name)
    vi = make_product(fx)
    ckpt = {k[len('resume_ckpt_'):]: v for k, v in fx.items() if k.startswith('resume_ckpt_')}
    params = vi.optimize(ckpt)
    tr = vi.trajectory
    assert tr['trials'] == fx['resume_traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['resume_traj_L0'])
    assert np.allclose(tr['elbo'], fx['resume_traj_elbo_out'], rtol=1e-8)
    assert np.allclose(params[0], fx['resume_final_vi_mu'], rtol=1e-6, atol=1e-9)
    assert np.allclose(params[2], fx['resume_final_hyper_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(vi.error_scaling, fx['resume_final_error_scaling'], rtol=1e-8)


def test_inputs_not_mutated_and_step_consistency():
    """tests/test.py:1529-1574: updates return new arrays; _optimize_step == _nat_grad_step."""
    fx = load_case('vischeme_unlinked_a1_s0_t0')
    vi = make_product(fx)
    np.random.seed(42)
    mu, delta, hyper = vi._initialize()
    copies = (mu.copy(), delta.copy(), hyper.copy())
    params = (mu, delta, hyper)
    g0 = np.copy(vi.nat_grad_vi_delta)
    new_params, new_L, d = vi._nat_grad_step(params, [1., 1., 1., 1., 1.], 2., None)
    for a, b in zip(params, copies):
        assert np.array_equal(a, b)
    assert new_L[0] == 1.
    assert vi.elbo(new_params) > vi.elbo(params)
    vi.nat_grad_vi_delta = g0
    opt_params, L2, elbo, run = vi._optimize_step(params, [1., 1., 1., 1., 1.], vi.elbo(params),
                                                  line_search_rate=2.)
    for a, b in zip(new_params, opt_params):
        assert np.allclose(a, b)
    assert np.isclose(elbo, vi.elbo(opt_params))


def test_line_search_from_huge_L():
    """tests/test.py:1499-1514: at L ~ L_MAX the step is tiny and mu barely moves."""
    from vilma_b200 import variational_inference as vin
    fx = load_case('vischeme_unlinked_a1_s0_t0')
    vi = make_product(fx)
    np.random.seed(42)
    params = vi._initialize()
    new_params, new_L, d = vi._nat_grad_step(params, [vin.L_MAX - 1, 1., 1.], 2., None)
    assert new_L[0] < vin.L_MAX - 1
    assert vi.elbo(new_params) > vi.elbo(params)
    assert np.allclose(params[0], new_params[0])


@pytest.mark.parametrize('symmetric', [1, 0])
def test_big_block_slabs_and_factor(symmetric):
    """Symmetric-packed blocks (ragged sizes, many panels/groups; blocks wider than VB_SYM_NMAX = 2816 are
    cut into 2 and 3 column slabs: 3000, 5000, 5700), full-storage blocks wider than one column slab
    (packing off) and tall factor blocks."""
    from vilma_b200.engine import DeviceContext, DeviceLD, set_option
    rng = np.random.default_rng(0)
    ctx = DeviceContext.get()
    sizes = [2500, 37, 1, 8, 515, 2816, 3000, 5000, 5700, 2817]
    mats = []
    for n in sizes:
        a = rng.standard_normal((n, n))
        mats.append(a + a.T)
    n3, r3 = 700, 150
    U = np.linalg.qr(rng.standard_normal((n3, r3)))[0]
    s = rng.uniform(0.5, 2.0, r3)
    tot = sum(sizes) + n3
    M = tot + 5
    perm = rng.permutation(M)[:tot]
    blocks = [{'n': n, 'kind': 'dense', 'R': R} for n, R in zip(sizes, mats)]
    blocks.append({'n': n3, 'kind': 'factor', 'U': U, 's': s})
    set_option('ld_symmetric', symmetric)
    try:
        ld = DeviceLD(ctx, M, blocks, perm)
    finally:
        set_option('ld_symmetric', 1)
    x = rng.standard_normal(M)
    y = ld.dot(x)
    ref = np.zeros(M)
    off = 0
    for n, R in zip(sizes, mats):
        idx = perm[off:off + n]
        ref[idx] = R @ x[idx]
        off += n
    idx = perm[off:]
    ref[idx] = U @ (s * (U.T @ x[idx]))
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-11 * np.abs(ref).max())
    dense = sum((4 * n * (n + 1) if symmetric else 8 * n * n) for n in sizes)
    assert ld.bytes == dense + 8 * n3 * r3          # the factor block is read once (U sqrt(s))
    # bit-reproducible across launches (dynamic scheduling must not change the summation order)
    assert np.array_equal(y, ld.dot(x))
    ld.close()


def test_widest_block_of_the_6m_snp_workload():
    """One symmetric-packed block of 13 827 rows -- the largest block of BASELINE configs[3] (bench.py
    --workload c4) -- is cut into five column slabs; against the dense product, next to a small block so that
    the finish kernel mixes single-slab and multi-slab records."""
    from vilma_b200.engine import DeviceContext, DeviceLD
    rng = np.random.default_rng(11)
    ctx = DeviceContext.get()
    n = 13827
    a = rng.standard_normal((n, n))
    big = a + a.T
    del a
    small = rng.standard_normal((40, 40))
    small = small + small.T
    M = n + 40 + 3
    perm = rng.permutation(M)[:n + 40]
    ld = DeviceLD(ctx, M, [{'n': 40, 'kind': 'dense', 'R': small}, {'n': n, 'kind': 'dense', 'R': big}], perm)
    assert ld.bytes == 4 * n * (n + 1) + 4 * 40 * 41
    x = rng.standard_normal(M)
    y = ld.dot(x)
    ref = np.zeros(M)
    ref[perm[:40]] = small @ x[perm[:40]]
    ref[perm[40:]] = big @ x[perm[40:]]
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max()), np.abs(y - ref).max()
    assert np.array_equal(y, ld.dot(x))
    ld.close()


@pytest.mark.parametrize('once', [1, 0])
def test_factor_blocks_read_once_and_two_pass(once):
    """Factor blocks R = U diag(s) U^T in both device forms -- read once as U sqrt(s) (n <= 2816: odd n, a
    single column, one column per chunk at n = 2816, several groups per block) and the two-pass form
    V' = diag(s) U^T, U (option off, and always above 2816 rows) -- mixed with packed dense blocks;
    against the dense product (reference LowRankMatrix.dot, matrix_structures.py:148-152)."""
    from vilma_b200._lib import VilmaB200Error
    from vilma_b200.engine import DeviceContext, DeviceLD, fac_nmax, set_option
    rng = np.random.default_rng(5)
    ctx = DeviceContext.get()
    shapes = [(700, 150), (37, 5), (1, 1), (2816, 40), (2815, 333), (513, 256), (3000, 60), (2, 1), (1201, 600)]
    blocks, refs = [], []
    for n, r in shapes:
        U = np.linalg.qr(rng.standard_normal((n, r)))[0]
        s = rng.uniform(0.0, 2.0, r)
        s[0] = 0.0                                   # a zero weight is fine in either form
        blocks.append({'n': n, 'kind': 'factor', 'U': np.ascontiguousarray(U), 's': s})
        refs.append((U * s) @ U.T)
    a = rng.standard_normal((300, 300))
    blocks.insert(3, {'n': 300, 'kind': 'dense', 'R': a + a.T})
    refs.insert(3, a + a.T)
    tot = sum(b['n'] for b in blocks)
    M = tot + 7
    perm = rng.permutation(M)[:tot]
    set_option('ld_factor_once', once)
    try:
        assert fac_nmax() == (2816 if once else 0)
        ld = DeviceLD(ctx, M, blocks, perm)
        bad = None
        if once:
            with pytest.raises(VilmaB200Error, match='negative'):
                bad = DeviceLD(ctx, 8, [{'n': 8, 'kind': 'factor', 'U': np.eye(8)[:, :2].copy(),
                                         's': np.array([1.0, -0.5])}], np.arange(8))
    finally:
        set_option('ld_factor_once', 1)
    assert bad is None
    x = rng.standard_normal(M)
    y = ld.dot(x)
    ref = np.zeros(M)
    off = 0
    expect = 0
    for b, R in zip(blocks, refs):
        n = b['n']
        idx = perm[off:off + n]
        ref[idx] = R @ x[idx]
        off += n
        if b['kind'] == 'dense':
            expect += 4 * n * (n + 1)
        else:
            r = b['U'].shape[1]
            expect += 8 * (n + (n & 1)) * r if once and n <= 2816 else 16 * n * r
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    missing = np.setdiff1d(np.arange(M), perm)
    assert np.all(y[missing] == 0)
    assert ld.bytes == expect
    assert np.array_equal(y, ld.dot(x))          # bit-reproducible under dynamic claiming
    ld.close()


def test_low_rank_blocks_never_carry_negative_weights():
    """The read-once factor form needs s >= 0.  LowRankMatrix keeps only s > 1e-12 max(s), exactly as the
    reference (matrix_structures.py:18, :119), so a caller-made block with a negative weight loses it before
    upload, and the device operator is the reference's."""
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    rng = np.random.default_rng(6)
    n, r = 400, 20
    U = np.linalg.qr(rng.standard_normal((n, r)))[0]
    s = rng.uniform(0.5, 2.0, r)
    s[3] = -0.7
    blk = LowRankMatrix(u=U, s=s, v=U.T.copy(), D=np.zeros(n))
    assert blk.s.shape == (r - 1,) and blk.s.min() > 0
    ld = BlockDiagonalMatrix([blk])
    x = rng.standard_normal(n)
    keep = s > 0
    ref = (U[:, keep] * s[keep]) @ (U[:, keep].T @ x)
    y = ld.dot(x)
    assert np.allclose(y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max()), np.abs(y - ref).max()
    assert ld.to_device().bytes == 8 * n * (r - 1)
    ld.release_device()


def test_fails_loudly_on_bad_input():
    from vilma_b200._lib import VilmaB200Error
    from vilma_b200.engine import DeviceContext, DeviceLD
    ctx = DeviceContext.get()
    with pytest.raises(VilmaB200Error):
        DeviceLD(ctx, 4, [{'n': 3, 'kind': 'dense', 'R': np.eye(3)}], np.array([0, 1, 1]))


def test_sym_and_full_storage_agree():
    """The symmetric-packed and the full dense stores are the same operator (to rounding)."""
    import torch
    from vilma_b200.engine import DeviceContext, DeviceLD, set_option
    rng = np.random.default_rng(9)
    sizes = [700, 64, 1500, 333]
    mats = []
    for n in sizes:
        a = rng.standard_normal((n, n))
        mats.append(a @ a.T / n)
    M = sum(sizes)
    blocks = [{'n': n, 'kind': 'dense', 'R': R} for n, R in zip(sizes, mats)]
    ctx = DeviceContext.get()
    packed = DeviceLD(ctx, M, blocks, np.arange(M))
    set_option('ld_symmetric', 0)
    try:
        full = DeviceLD(ctx, M, blocks, np.arange(M))
    finally:
        set_option('ld_symmetric', 1)
    x = rng.standard_normal(M)
    yp, yf = packed.dot(x), full.dot(x)
    assert packed.bytes < 0.55 * full.bytes
    assert np.allclose(yp, yf, rtol=1e-13, atol=1e-13 * np.abs(yf).max())
    packed.close()
    full.close()


def test_checkpointing_does_not_disturb_speculation(tmp_path):
    """Periodic checkpoints download the state between iterations, which invalidates the trial
    queued speculatively behind the previous iteration: the fit must fall back and reproduce the
    same trajectory (syn_p5 has no error-scaling step, so speculation is active)."""
    fx = load_case('syn_p5')
    vi = make_product(fx, checkpoint=True, checkpoint_freq=3, output=str(tmp_path / 'ck'))
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=0)
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=1e-9)
    import os
    files = sorted(os.listdir(tmp_path))
    assert files == sorted('ck-checkpoint.%d.npz' % i for i in range(0, len(tr['elbo']), 3))
    first = np.load(tmp_path / 'ck-checkpoint.0.npz')
    assert np.allclose(first['vi_mu'], fx['init_vi_mu'], rtol=1e-7, atol=1e-12)


def test_closed_form_covariances():
    """The reference's hand-computed covariances for its unit fixture (tests/test.py:1297-1352) from the
    device kernels: S_ki = diag(1/2, 4/5), diag(2/3, 4/3); log|S| = log(2/5), log(8/9);
    tr(Prec_k S_ki) = 13/10, 1."""
    fx = load_case('vischeme_linked_a2_s0_t1')
    vi = make_product(fx)
    M = fx['betas'].shape[1]
    true_sigma = np.zeros((2, 2, 2, M))
    true_sigma[0, 0, 0], true_sigma[0, 1, 1] = 1 / 2, 4 / 5
    true_sigma[1, 0, 0], true_sigma[1, 1, 1] = 2 / 3, 4 / 3
    assert np.allclose(vi.vi_sigma, true_sigma, rtol=1e-14, atol=1e-15)
    true_nat = np.zeros((2, 2, 2, M))
    true_nat[0, 0, 0], true_nat[0, 1, 1] = -1, -5 / 8
    true_nat[1, 0, 0], true_nat[1, 1, 1] = -3 / 4, -3 / 8
    assert np.allclose(vi.nat_sigma, true_nat)
    true_ld = np.array([[np.log(2 / 5)] * M, [np.log(8 / 9)] * M])
    assert np.allclose(vi.vi_sigma_log_det, true_ld)
    true_matches = np.stack([np.full(M, 1 / 2 + 4 / 5), np.full(M, 1 / 3 + 2 / 3)], axis=1)
    assert np.allclose(vi.vi_sigma_matches, true_matches)
    assert np.allclose(vi.sigma_summary, np.array([0., 2 * np.log(2)]) - true_ld.T + true_matches)
    assert vi.nat_grad_vi_delta is None
