"""Larger-than-fixture checks on the GPU (BASELINE.json configs[1] shapes).

* a ~25k-SNP slice of the benchmark workload fitted by the product and by the oracle: same
  line-search decisions, ELBO trajectory to 1e-8, posterior means to 1e-6;
* size-independent properties at a 240k-SNP slice (340 blocks): symmetric-packed and full LD storage
  give the same operator, the operator is symmetric (x.Ry == y.Rx) and linear, every accepted update
  raises the ELBO within the reference's tolerance, the tracked ELBO equals a recomputed one, the
  mixture weights stay normalised.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_medium_fit_matches_oracle():
    import bench
    from vilma_b200.dist import SingleComm
    n_blocks = 36
    M_total = int(round(bench.layout()[1][:n_blocks].sum() / (1 - bench.MISSING_FRAC)))
    # product: blocks generated on the device (same generator as bench.py)
    import torch
    torch.cuda.set_device(0)
    # the oracle sample uses the first n_blocks of the C2 layout; build the product on exactly those
    vi_o, M_o, _ = bench.build_cpu_sample(n_blocks)
    M_ld, n_all, starts = bench.layout()
    n = n_all[:n_blocks]
    vi_p = _product_from_oracle_inputs(vi_o, n)
    np.random.seed(42)
    p_o = vi_o._initialize()
    its = 8
    vi_o.num_its = its
    traj = {}
    np.random.seed(42)
    res_o = vi_o.optimize(None, trajectory=traj)
    vi_p.num_its = its
    np.random.seed(42)
    res_p = vi_p.optimize(None)
    assert vi_p.trajectory['trials'] == traj['trials']
    assert np.array_equal(np.array(vi_p.trajectory['L0']), np.array(traj['L0']))
    assert np.allclose(vi_p.trajectory['elbo'], traj['elbo_out'], rtol=1e-8, atol=0)
    pm_o = vi_o.real_posterior_mean(*res_o)
    pm_p = vi_p.real_posterior_mean(*res_p)
    assert np.allclose(pm_p, pm_o, rtol=1e-6, atol=1e-9)
    assert np.allclose(res_p[2], res_o[2], rtol=1e-6, atol=1e-12)


def _product_from_oracle_inputs(vi_o, n):
    """Product MultiPopVI over the SAME LD (dense reconstruction of the oracle's blocks) and the
    oracle's own set-up values, so that only the fitting loop is under test."""
    from vilma_b200.engine import DeviceContext, DeviceLD
    from vilma_b200.variational_inference import DeviceBlockDiagonalMatrix, MultiPopVI
    ld_o = vi_o.ld_mats[0]
    M = vi_o.num_loci
    ctx = DeviceContext(0)
    blocks = [{'n': int(b.shape[0]), 'kind': 'dense', 'R': (b.u * b.s) @ b.v} for b in ld_o.matrices]
    ld = DeviceLD(ctx, M, blocks, ld_o.perm[:int(ld_o.starts[-1])])
    pre = dict(ld_diags=vi_o.ld_diags, adj_marginal_effects=vi_o.adj_marginal_effects,
               chi_stat=vi_o.chi_stat, ld_ranks=vi_o.ld_ranks, inverse_betas=vi_o.inverse_betas)
    covs = [np.linalg.inv(vi_o.mixture_prec[k, :, :, 0]) for k in range(vi_o.num_mix)]
    return MultiPopVI(marginal_effects=vi_o.marginal_effects, std_errs=vi_o.std_errs,
                      ld_mats=[DeviceBlockDiagonalMatrix(ld, (M, M))], mixture_covs=covs,
                      annotations=np.ones((M, 1)), checkpoint=False, checkpoint_freq=-1, output='t',
                      scaled=False, scale_se=False, gwas_N=np.array([3e5]), init_hg=np.array([0.3]),
                      num_its=10, comm=None, device=0, precomputed=pre, context=ctx)


def test_large_properties():
    import torch
    import bench
    from vilma_b200.dist import SingleComm
    from vilma_b200.engine import DeviceLD, set_option
    vi, ctx, info = bench.build_gpu_problem(SingleComm(), 0, M_total=240_000, n_blocks=340, num_its=6)
    M = info['M']
    ld = vi.ld_mats[0].device_ld
    gen = torch.Generator(device='cuda')
    gen.manual_seed(3)
    x = torch.randn(M, generator=gen, device='cuda', dtype=torch.float64)
    y = torch.randn(M, generator=gen, device='cuda', dtype=torch.float64)
    Rx, Ry, Rxy = (torch.empty_like(x) for _ in range(3))
    ld.dot_device(x, Rx)
    ld.dot_device(y, Ry)
    ld.dot_device(2.0 * x - 3.0 * y, Rxy)
    # symmetry and linearity of the operator
    a, b = float(y @ Rx), float(x @ Ry)
    assert abs(a - b) <= 1e-11 * max(abs(a), 1.0)
    assert torch.allclose(Rxy, 2.0 * Rx - 3.0 * Ry, rtol=1e-11, atol=1e-9)
    assert torch.all(Rx[info['M_ld']:] == 0)          # SNPs without LD give zero rows
    # the fit: monotone accepted updates, tracked == recomputed ELBO, normalised weights
    np.random.seed(42)
    params = vi.optimize(None)
    el = np.array(vi.trajectory['elbo'])
    assert np.all(np.diff(el) >= -1e-6 * np.abs(el[:-1]) - 1e-6)
    assert np.isclose(vi.elbo(params), el[-1], rtol=1e-10)
    assert np.allclose(params[1].sum(axis=1), 1.0, atol=1e-12)
    assert np.allclose(params[2].sum(axis=1), 1.0, atol=1e-12)
    assert np.all(params[1] >= 1e-100)
    pm = vi.real_posterior_mean(*params)
    assert np.all(pm[:, info['M_ld']:] == 0.0)         # no LD, BETA 0 -> posterior mean 0


def test_medium_multi_cohort_fit_matches_oracle():
    """A ~4k-SNP slice of the three-cohort workload (bench.py --workload c3: low-rank panels, 87
    components): device _initialize + the tile kernel against the oracle on identical inputs --
    same line-search decisions, ELBO trajectory, posterior means."""
    import bench
    from vilma_b200.engine import DeviceContext, DeviceLD
    from vilma_b200.variational_inference import DeviceBlockDiagonalMatrix, MultiPopVI
    vi_o, M, _ = bench.build_cpu_sample_multi(bench.WORKLOADS['c3'], 6)
    P = vi_o.num_pops
    ctx = DeviceContext(0)
    lds = []
    for ld_o in vi_o.ld_mats:
        blocks = [{'n': int(b.shape[0]), 'kind': 'dense', 'R': (b.u * b.s) @ b.v} for b in ld_o.matrices]
        lds.append(DeviceLD(ctx, M, blocks, ld_o.perm[:int(ld_o.starts[-1])]))
    pre = dict(ld_diags=vi_o.ld_diags, adj_marginal_effects=vi_o.adj_marginal_effects,
               chi_stat=vi_o.chi_stat, ld_ranks=vi_o.ld_ranks, inverse_betas=vi_o.inverse_betas)
    covs = [np.linalg.inv(vi_o.mixture_prec[k, :, :, 0]) for k in range(vi_o.num_mix)]
    vi_p = MultiPopVI(marginal_effects=vi_o.marginal_effects, std_errs=vi_o.std_errs,
                      ld_mats=[DeviceBlockDiagonalMatrix(ld, (M, M)) for ld in lds], mixture_covs=covs,
                      annotations=np.ones((M, 1)), checkpoint=False, checkpoint_freq=-1, output='t',
                      scaled=False, scale_se=False, gwas_N=np.array(bench.WORKLOADS['c3']['N']),
                      init_hg=np.full(P, bench.INIT_HG), num_its=6, comm=None, device=0,
                      precomputed=pre, context=ctx)
    vi_p.init_on_device = True
    vi_o.num_its = 6
    traj = {}
    np.random.seed(42)
    res_o = vi_o.optimize(None, trajectory=traj)
    np.random.seed(42)
    res_p = vi_p.optimize(None)
    assert vi_p.trajectory['trials'] == traj['trials']
    assert np.array_equal(np.array(vi_p.trajectory['L0']), np.array(traj['L0']))
    el = np.array(traj['elbo_out'])
    assert np.allclose(vi_p.trajectory['elbo'], el, rtol=1e-8, atol=1e-13 * np.abs(el).max())
    assert np.allclose(vi_p.real_posterior_mean(*res_p), vi_o.real_posterior_mean(*res_o), rtol=1e-6, atol=1e-9)
    assert np.allclose(res_p[2], res_o[2], rtol=1e-6, atol=1e-12)
