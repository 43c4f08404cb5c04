"""Loader / grid known-answer tests on the reference's own test data (packed in the CLI goldens).

Mirrors /root/reference/tests/test.py:486-706 (loaders) and the grid construction of
vi_options.py:196-337; LD operators are checked through a dense reconstruction here (the
`.dot` itself is a GPU test).
"""
import os
import numpy as np
import pytest

from _cli import materialize
from _fixtures import load_case


@pytest.fixture
def data(tmp_path):
    fx = load_case('cli_multi')
    materialize(fx, str(tmp_path))
    return fx, tmp_path


def dense(ld):
    n = ld.shape[0]
    out = np.zeros((n, n))
    for b, m in enumerate(ld.matrices):
        idx = ld.perm[ld.starts[b]:ld.starts[b + 1]]
        out[np.ix_(idx, idx)] = (m.u * m.s) @ m.v + np.diag(m.D)
    return out


def true_ld(zero):
    t = np.eye(13)
    t[0, 2] = t[2, 0] = -1
    for i in zero:
        t[i, i] = 0
    return t


def test_variants_annotations(data):
    from vilma_b200 import load
    fx, d = data
    variants = load.load_variant_list(str(d / 'good_variants.tsv'))
    assert len(variants) == 13 and list(variants.columns) == ['ID', 'A1', 'A2']
    null_ann, deny = load.load_annotations(None, variants)
    assert null_ann.shape == (13, 1) and np.allclose(null_ann, 1) and deny == []
    ann, deny = load.load_annotations(str(d / 'good_annotations.tsv'), variants)
    assert ann.shape == (13, 6)
    assert np.all(ann.sum(axis=1) == 1)
    assert np.all(ann.sum(axis=0)[1:] == 2) and ann.sum(axis=0)[0] == 3
    assert deny == [12]


def test_sumstats(data):
    from vilma_b200 import load
    fx, d = data
    variants = load.load_variant_list(str(d / 'good_variants.tsv'))
    stats, deny = load.load_sumstats(str(d / 'good_sumstats_beta.tsv'), variants)
    assert set(deny) == {10, 11, 12} and len(stats) == 13
    assert np.all(stats.BETA.iloc[0:10] == np.arange(10)) and np.all(stats.BETA.iloc[10:] == 0)
    assert np.all(stats.SE.iloc[0:10] == np.arange(10) + 1) and np.all(stats.SE.iloc[10:] == 1)
    stats, deny = load.load_sumstats(str(d / 'good_sumstats_flip.tsv'), variants)
    assert set(deny) == {0, 10, 11, 12}
    assert np.all(stats.BETA.iloc[0:10] == -np.arange(10))
    assert np.all(stats.SE.iloc[0:10] == np.arange(10) + 1)
    bad = d / 'bad.tsv'
    bad.write_text('ID A1 A2 BETA\nx A C 1\n')
    with pytest.raises(ValueError):
        load.load_sumstats(str(bad), variants)


@pytest.mark.parametrize('manifest', ['ld_manifest.tsv', 'ld_manifest_svd.tsv'])
def test_ld_schema(data, manifest):
    from vilma_b200 import load
    fx, d = data
    variants = load.load_variant_list(str(d / 'good_variants.tsv'))
    ld, missing = load.load_ld_from_schema(str(d / manifest), variants, [], 1., False)
    assert sorted(missing) == [5, 12]
    assert np.allclose(dense(ld), true_ld([5, 12]))
    assert np.array_equal(np.sort(ld.perm), np.arange(13))
    ld, missing = load.load_ld_from_schema(str(d / manifest), variants, [3, 4, 5], 1., False)
    assert sorted(missing) == [3, 4, 5, 12]
    assert np.allclose(dense(ld), true_ld([3, 4, 5, 12]))
    # setup-only host operators
    assert ld.get_rank() == np.linalg.matrix_rank(true_ld([3, 4, 5, 12]))
    assert np.allclose(ld.diag(), np.diag(true_ld([3, 4, 5, 12])))
    e = np.zeros(13); e[12] = 1
    assert np.allclose(ld.inverse.dot(e), 0)


def test_mixture_grid_matches_reference_pickles(data):
    from vilma_b200 import load, vi_options
    fx, d = data
    variants = load.load_variant_list(str(d / 'good_variants.tsv'))
    b1, _ = load.load_sumstats(str(d / 'good_sumstats_beta.tsv'), variants)
    b2, _ = load.load_sumstats(str(d / 'good_sumstats_flip.tsv'), variants)
    # single cohort, -K 80: the reference's committed copy_vilma_run.covariance.pkl
    one = load_case('cli_fit')
    betas = np.array(b1.BETA)[None]
    ses = np.array(b1.SE)[None]
    mins, maxes = vi_options._grid_range(betas, ses, False)
    grid = vi_options._make_simple(1, 80, mins, maxes)
    assert np.allclose(np.array(grid), one['gold_covariance'], rtol=1e-12, atol=0)
    # two cohorts: RNG order matters (seed 7, --stderrscale 1.0,1.5, -K 3)
    betas = np.stack([np.array(b1.BETA), np.array(b2.BETA)])
    ses = np.stack([np.array(b1.SE), np.array(b2.SE) * 1.5])
    np.random.seed(7)
    mins, maxes = vi_options._grid_range(betas, ses, False)
    grid = vi_options._make_simple(2, 3, mins, maxes)
    assert np.allclose(np.array(grid), fx['run_covariance'], rtol=1e-12, atol=0)


def test_cli_fit_flag_surface_matches_reference():
    """`vilma fit` flags, defaults, types and required-ness, against the table recorded from the
    reference's own parser (vi_options.py:9-84; tests/golden/cli_fit_flags.json, written by
    introspecting the unmodified reference's argparse in the build container)."""
    import argparse
    import json
    from vilma_b200 import vi_options
    top = argparse.ArgumentParser(prog='vilma')
    parser = vi_options.args(top.add_subparsers())
    ours = {}
    for a in parser._actions:
        if not a.option_strings or a.dest == 'help':
            continue
        ours[a.dest] = dict(flags=sorted(a.option_strings), default=a.default, required=bool(a.required),
                            nargs=a.nargs, type=getattr(a.type, '__name__', None), action=type(a).__name__)
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                                      'cli_fit_flags.json')))
    assert ours == ref


def test_sim_draws_match_reference_loop():
    """vilma_b200.sim vectorises sim.sim_components (one np.random.choice per SNP in the reference,
    sim.py:73-96) and the 'ip,ik,kqp->qi' einsum (:99-137): same seed -> the same draws, bit for bit."""
    from oracle import ref_loader
    if not ref_loader.available()[0]:
        pytest.skip('oracle/_ref is not installed')
    import importlib
    ref_loader.import_reference()
    ref_sim = importlib.import_module('vilma.sim')
    from vilma_b200 import sim
    rng = np.random.default_rng(2)
    M, A, K, P = 3000, 3, 5, 2
    ann = np.zeros((M, A))
    ann[np.arange(M), rng.integers(0, A, M)] = 1
    w = rng.random((A, K))
    w /= w.sum(axis=1, keepdims=True)
    covs = []
    for _ in range(K):
        a = rng.standard_normal((P, P))
        covs.append(a @ a.T + 0.1 * np.eye(P))
    covs = np.array(covs)
    np.random.seed(5)
    want_c = ref_sim.sim_components(ann, w)
    np.random.seed(5)
    got_c = sim.sim_components(ann, w)
    assert np.array_equal(want_c, got_c)
    np.random.seed(6)
    want_e = ref_sim.sim_true_effects(ann, w, covs)
    np.random.seed(6)
    got_e = sim.sim_true_effects(ann, w, covs)
    assert np.allclose(want_e, got_e, rtol=1e-14, atol=0)
    # and the stream is left in the same place
    assert np.random.random_sample() == (np.random.seed(6), ref_sim.sim_true_effects(ann, w, covs),
                                         np.random.random_sample())[2]
