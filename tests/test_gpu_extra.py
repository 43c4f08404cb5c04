"""Parity on the fixtures of tests/golden/extra/: the reference's DEFAULT two-cohort grid (-K 12 -> 582
components: the tile kernel at 16 warps per tile) and 4 / 6 cohorts (the remaining P x P template
instantiations).  Same checks and tolerances as test_gpu_parity.py.

The oracle is pinned on the same fixtures in tests/test_oracle_golden.py.
"""
import numpy as np
import pytest

from _fixtures import EXTRA_CASES, build_ld, load_case, vi_kwargs

pytestmark = pytest.mark.gpu


def make_product(fx):
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    return MultiPopVI(ld_mats=build_ld(fx, LowRankMatrix, BlockDiagonalMatrix), **vi_kwargs(fx))


def mu_atol(fx, key):
    return max(1e-12, 1e-13 * np.abs(fx[key]).max())


@pytest.mark.parametrize('name', EXTRA_CASES)
def test_state_eval_and_device_init(name):
    fx = load_case(name)
    vi = make_product(fx)
    params = (fx['init_vi_mu'], fx['init_vi_delta'], fx['init_hyper_delta'])
    vi._set_state(params)
    assert np.isclose(vi.elbo(params), float(fx['init_elbo']), rtol=1e-11)
    assert np.allclose(vi.real_posterior_mean(*params), fx['init_post_mean'], rtol=1e-10, atol=1e-14)
    for on_device in (False, True):
        vi.init_on_device = on_device
        np.random.seed(int(fx['seed']))
        mu, delta, hyper = vi._initialize()
        assert np.allclose(mu, fx['init_vi_mu'], rtol=1e-7, atol=mu_atol(fx, 'init_vi_mu'))
        assert np.allclose(delta, fx['init_vi_delta'], rtol=1e-7, atol=1e-300)
        assert np.allclose(hyper, fx['init_hyper_delta'], rtol=1e-9)


@pytest.mark.parametrize('name', EXTRA_CASES)
def test_trajectory(name):
    fx = load_case(name)
    vi = make_product(fx)
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    # accumulated ELBO: floor relative to the trajectory's scale (see tests/test_oracle_golden.py)
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8,
                       atol=1e-13 * np.abs(fx['traj_elbo_out']).max())
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=max(1e-9, mu_atol(fx, 'final_vi_mu')))
    assert np.allclose(params[1], fx['final_vi_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(params[2], fx['final_hyper_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(vi.real_posterior_mean(*params), fx['final_post_mean'], rtol=1e-6, atol=1e-9)
    if 'final_vi_sigma' in fx:      # slim fixtures (K = 256) leave the [K,P,P,M] array out
        assert np.allclose(vi.vi_sigma, fx['final_vi_sigma'], rtol=1e-8)
