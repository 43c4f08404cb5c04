"""End-to-end `vilma fit` through the drop-in CLI against the reference's golden outputs.

Mirrors /root/reference/tests/test.py:2161-2197 (test_cli_fit) and the example scripts.
"""
import os
import pickle

import numpy as np
import pytest

from _cli import frames_close, materialize, read_tsv
from _fixtures import load_case

pytestmark = pytest.mark.gpu


def run_cli(argv):
    from vilma_b200 import frontend
    assert frontend.main(argv) == 0


def test_cli_fit(tmp_path):
    fx = load_case('cli_fit')
    d = materialize(fx, str(tmp_path))
    out = os.path.join(d, 'vilma_run')
    run_cli(['fit', '--ld-schema', os.path.join(d, 'ld_manifest.tsv'),
             '--sumstats', os.path.join(d, 'good_sumstats_beta.tsv'), '--output', out,
             '-K', '80', '--ldthresh', '0.8', '--init-hg', '0.2', '--samplesizes', '10e3',
             '--names', 'test_cohort', '--learn-scaling',
             '--extract', os.path.join(d, 'good_variants.tsv')])
    cli = np.load(out + '.npz')
    gold_keys = [k[len('gold_npz_'):] for k in fx if k.startswith('gold_npz_')]
    assert sorted(cli.files) == sorted(gold_keys)
    for k in gold_keys:
        assert cli[k].shape == fx['gold_npz_' + k].shape and cli[k].dtype == fx['gold_npz_' + k].dtype
        assert np.allclose(cli[k], fx['gold_npz_' + k]), k                  # the reference's own check
        assert np.allclose(cli[k], fx['gold_npz_' + k], rtol=1e-6, atol=1e-9), k
    with open(out + '.covariance.pkl', 'rb') as fh:
        covs = pickle.load(fh)
    assert np.allclose(covs, fx['gold_covariance'][None])
    assert frames_close(read_tsv(fx['gold_estimates_tsv']), read_tsv(open(out + '.estimates.tsv').read()),
                        rtol=1e-6, atol=1e-9)


def test_cli_example_and_resume(tmp_path):
    fx = load_case('cli_example')
    d = materialize(fx, str(tmp_path))
    common = ['--sumstats', os.path.join(d, 'example_data/example_gwas_sumstats.txt'),
              '--ld-schema', os.path.join(d, 'ld_mat/example_schema.schema'), '--seed', '42',
              '-K', '81', '--init-hg', '0.2', '--samplesizes', '300e3', '--names', 'ukbb',
              '--learn-scaling', '--extract', os.path.join(d, 'keep_variants.txt')]
    out = os.path.join(d, 'example_vilma_run')
    run_cli(['fit', '--output', out] + common)
    assert frames_close(read_tsv(fx['gold_estimates_tsv']), read_tsv(open(out + '.estimates.tsv').read()),
                        rtol=1e-6, atol=1e-9)
    out2 = os.path.join(d, 'checkpoint_example_vilma_run')
    run_cli(['fit', '--output', out2, '--load-checkpoint', out + '.npz', out + '.covariance.pkl'] + common)
    cli = np.load(out2 + '.npz')
    for k in cli.files:
        tol = 1e-6 if k != 'error_scaling' else 1e-5
        assert np.allclose(cli[k], fx['resume_gold_npz_' + k], rtol=tol, atol=1e-9), k
    assert frames_close(read_tsv(fx['resume_gold_estimates_tsv']),
                        read_tsv(open(out2 + '.estimates.tsv').read()), rtol=1e-6, atol=1e-9)


def test_cli_two_cohorts_annotations_checkpoints(tmp_path):
    fx = load_case('cli_multi')
    d = materialize(fx, str(tmp_path))
    out = os.path.join(d, 'multi_run')
    j = lambda f: os.path.join(d, f)
    run_cli(['fit', '--ld-schema', j('ld_manifest.tsv') + ',' + j('ld_manifest_svd.tsv'),
             '--sumstats', j('good_sumstats_beta.tsv') + ',' + j('good_sumstats_flip.tsv'),
             '--annotations', j('good_annotations.tsv'), '--output', out, '-K', '3',
             '--ldthresh', '0.9', '--init-hg', '0.2,0.3', '--samplesizes', '10e3,20e3',
             '--names', 'a,b', '--stderrscale', '1.0,1.5', '--learn-scaling', '--seed', '7',
             '--checkpoint-freq', '4', '--num-its', '30', '--extract', j('good_variants.tsv')])
    cli = np.load(out + '.npz')
    for k in cli.files:
        assert np.allclose(cli[k], fx['run_npz_' + k], rtol=1e-6, atol=1e-9), k
    assert frames_close(read_tsv(fx['run_estimates_tsv']), read_tsv(open(out + '.estimates.tsv').read()),
                        rtol=1e-6, atol=1e-9)
    ckpts = sorted(f for f in os.listdir(d) if f.startswith('multi_run-checkpoint.'))
    assert ckpts == [str(c) for c in fx['checkpoint_files']]
    for f in ckpts:
        z = np.load(os.path.join(d, f))
        for k in z.files:
            assert np.allclose(z[k], fx['ckpt_%s_%s' % (f, k)], rtol=1e-6, atol=1e-9), (f, k)


def test_loader_ld_dot_known_answer(tmp_path):
    """tests/test.py:595-657: the loaded operator applied on the GPU equals the known matrix."""
    from vilma_b200 import load
    fx = load_case('cli_multi')
    d = materialize(fx, str(tmp_path))
    variants = load.load_variant_list(os.path.join(d, 'good_variants.tsv'))
    truth = np.eye(13); truth[0, 2] = truth[2, 0] = -1; truth[5, 5] = truth[12, 12] = 0
    for manifest in ('ld_manifest.tsv', 'ld_manifest_svd.tsv'):
        ld, missing = load.load_ld_from_schema(os.path.join(d, manifest), variants, [], 1., False)
        v = np.random.random(13)
        assert np.allclose(ld.dot(v), truth.dot(v))


def test_cli_sim_matches_reference_golden(tmp_path):
    """`vilma sim` against the reference's own golden output
    (tests/test.py:2200-2246, copy_vilma_sim_run.simpop1.simgwas.tsv, --seed 143), with the weights
    given as a .npy matrix and as a fitted model's .npz; both LD products run on the GPU."""
    fx = load_case('cli_sim')
    d = materialize(fx, str(tmp_path))
    with open(os.path.join(d, 'covs.pkl'), 'wb') as fh:
        pickle.dump([list(fx['covariance'])], fh)
    np.savez(os.path.join(d, 'sim_weights_model.npz'), hyper_delta=fx['weights_npz_hyper_delta'])
    gold = read_tsv(fx['gold_simgwas_tsv'])
    for weights, out in (('sim_weights.npy', 'run_npy'), ('sim_weights_model.npz', 'run_npz')):
        run_cli(['sim', '--ld-schema', os.path.join(d, 'ld_manifest.tsv'),
                 '--sumstats', os.path.join(d, 'good_sumstats_beta.tsv'),
                 '--annotations', os.path.join(d, 'good_annotations.tsv'),
                 '--covariance', os.path.join(d, 'covs.pkl'), '--weights', os.path.join(d, weights),
                 '--output', os.path.join(d, out), '--names', 'simpop1', '--seed', '143'])
        got = read_tsv(open(os.path.join(d, out + '.simpop1.simgwas.tsv')).read())
        assert frames_close(gold, got)             # the reference's own tolerance (check_data_frame)
        assert frames_close(gold, got, rtol=1e-9, atol=1e-12)
