"""CPU tests of the product's host logic: control flow, sharding, rank plumbing, C ABI surface.

The per-SNP / LD arithmetic is supplied by a TEST-ONLY NumPy engine (tests/_np_engine.py, built
on the oracle) through the ``engine_factory`` hook, so these tests exercise exactly the Python
that drives the GPU in production, without a GPU.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from _fixtures import VI_CASES, build_ld, load_case, vi_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def precomputed(fx):
    return dict(ld_diags=fx['pre_ld_diags'], adj_marginal_effects=fx['pre_adj_marginal_effects'],
                chi_stat=fx['pre_chi_stat'], ld_ranks=fx['pre_ld_ranks'],
                inverse_betas=fx['pre_inverse_betas'])


def make_host_vi(fx, comm=None):
    from _np_engine import NumpyShardEngine
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    return MultiPopVI(ld_mats=build_ld(fx, LowRankMatrix, BlockDiagonalMatrix),
                      precomputed=precomputed(fx), engine_factory=NumpyShardEngine, comm=comm,
                      **vi_kwargs(fx))


def check_trajectory(vi, params, fx):
    tr = vi.trajectory
    assert tr['trials'] == fx['traj_trials'].tolist()
    assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
    assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-9, atol=0)
    assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6, atol=1e-9)
    assert np.allclose(params[1], fx['final_vi_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(params[2], fx['final_hyper_delta'], rtol=1e-6, atol=1e-12)
    assert np.allclose(vi.error_scaling, fx['final_error_scaling'], rtol=1e-8)


@pytest.mark.parametrize('name', ['vischeme_linked_a2_s0_t1', 'syn_p1_dense', 'syn_p2_lowrank', 'syn_p3'])
def test_host_loop_single_rank(name):
    """The cached-objective control flow reproduces the reference's decision sequence."""
    fx = load_case(name)
    vi = make_host_vi(fx)
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    check_trajectory(vi, params, fx)
    # one evaluation per distinct state: trials + hyper (+ tau) per iteration, + the initial ones
    its = len(fx['traj_trials'])
    assert vi.n_evals <= int(fx['traj_trials'].sum()) + 2 * its + 2


def test_partition_properties():
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.partition import host_block_lists, local_blocks, partition_snps
    fx = load_case('syn_p2_lowrank')
    lds = build_ld(fx, LowRankMatrix, BlockDiagonalMatrix)
    M = fx['betas'].shape[1]
    for world in (1, 2, 3, 4):
        parts = partition_snps(host_block_lists(lds), M, world)
        allsnps = np.sort(np.concatenate(parts))
        assert np.array_equal(allsnps, np.arange(M))            # a partition of the SNPs
        for snps in parts:
            for ld in lds:
                ids, perm_local = local_blocks(ld, snps, M)      # raises if a block is split
                assert len(perm_local) == sum(ld.matrices[b].shape[0] for b in ids)
                assert len(np.unique(perm_local)) == len(perm_local)
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= max(80, M // world)    # roughly balanced


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import torch.distributed as dist
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
from vilma_b200.dist import TorchComm
from test_host_logic import make_host_vi, check_trajectory
from _fixtures import load_case
fx = load_case({name!r})
vi = make_host_vi(fx, comm=TorchComm())
assert 0 < len(vi._snps) < fx['betas'].shape[1]
np.random.seed(int(fx['seed']))
params = vi.optimize(None)
check_trajectory(vi, params, fx)
pm = vi.real_posterior_mean(*params)
assert np.allclose(pm, fx['final_post_mean'], rtol=1e-6, atol=1e-9)
dist.barrier(); dist.destroy_process_group()
print('rank', sys.argv[1], 'ok')
'''


@pytest.mark.parametrize('name,no_shm', [('syn_p1_dense', '0'), ('syn_p2_lowrank', '0'), ('syn_p1_dense', '1')])
def test_two_ranks_gloo(name, no_shm, tmp_path):
    """world_size=2 over gloo: sharded SNPs + all-reduced statistics give the single-process
    trajectory (same decisions on every rank).  The fitted parameters come back through the node-shared
    mapping, or -- VILMA_B200_NO_SHM=1, as on a host whose /dev/shm is too small -- through all-gathers."""
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER.format(root=ROOT, port=port, name=name))
    env = dict(os.environ, VILMA_B200_NO_SHM=no_shm)
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, 'rank %d failed:\n%s' % (r, o)
        assert 'rank %d ok' % r in o


_POOL_WORKER = r"""
import gc, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
from vilma_b200.dist import TorchComm, _SHM_POOL
comm = TorchComm()
rank = comm.rank
shapes = [(3, 2, 1000), (1000, 3)]
(a, b), e1 = comm.shared_arrays(shapes)
mine = slice(0, 500) if rank == 0 else slice(500, 1000)
a[..., mine] = rank + 1.0
b[mine] = 10.0 * (rank + 1)
comm.barrier()
assert np.all(a[..., :500] == 1.0) and np.all(a[..., 500:] == 2.0)          # the other rank's half is visible
assert np.all(b[:500] == 10.0) and np.all(b[500:] == 20.0)
addr1 = e1['address']
# still referenced (here on rank 0 only): a second request must get a NEW mapping on both ranks
keep = a[0] if rank == 0 else None
del a, b
(c, d), e2 = comm.shared_arrays(shapes)
assert e2 is not e1 and len(_SHM_POOL) == 2
del c, d, keep
gc.collect()
# everything dropped everywhere: the first free mapping of that size is handed out again
(f, g), e3 = comm.shared_arrays(shapes)
assert e3 is e1 and e3['address'] == addr1 and len(_SHM_POOL) == 2
released = []
e2['release'] = lambda: released.append(1)
del f, g
# another size: the unreferenced mappings are released (hook called) and unmapped, one new mapping remains
(h,), e4 = comm.shared_arrays([(7, 11)])
assert len(_SHM_POOL) == 1 and _SHM_POOL[0] is e4 and released == [1]
h[:] = rank
comm.barrier()
# a second communicator object shares the process-wide pool
(k,), e5 = TorchComm().shared_arrays([(7, 11)])
assert e5 is not e4 and len(_SHM_POOL) == 2          # h is still referenced
print('rank %d ok' % rank)
dist.barrier()
dist.destroy_process_group()
"""


def test_shared_result_mappings_are_pooled_consistently(tmp_path):
    """TorchComm.shared_arrays over gloo, world_size=2: one node-shared mapping every rank writes its part
    of; a mapping is reused only when EVERY rank has dropped its arrays; a request for another size releases
    the unreferenced ones (through the hook a caller installed when it page-locked them) and the pools of
    the two ranks keep the same entries throughout (they vote with one all-reduce per request)."""
    port = 31500 + (os.getpid() % 2000)
    script = tmp_path / 'pool_worker.py'
    script.write_text(_POOL_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, 'rank %d failed:\n%s' % (r, o)
        assert 'rank %d ok' % r in o


def test_storage_choice_by_bytes_streamed():
    """engine.choose_storage picks, per LD block, the device form that streams the fewest bytes per mat-vec:
    packed dense 4 n (n+1), read-once factor 8 n_pad r (n <= vb_ld_fac_nmax), two-pass factor 16 n r."""
    from vilma_b200.engine import choose_storage, dense_bytes, fac_nmax, factor_bytes, sym_nmax
    assert fac_nmax() == 2816 and sym_nmax() == 65528
    assert dense_bytes(706) == 4 * 706 * 707 and dense_bytes(70000) == 8 * 70000 ** 2
    assert factor_bytes(706, 211) == 8 * 706 * 211 and factor_bytes(705, 10) == 8 * 706 * 10
    assert factor_bytes(3000, 100) == 16 * 3000 * 100
    # --ldthresh 0.99-like panels (r = 0.3 n): factors, read once (round 1 stored them dense: 16 n r > 4 n (n+1))
    assert choose_storage(706, 211) == 'factor'
    # the break-even of the read-once form is r = (n + 1) / 2, of the two-pass form r = (n + 1) / 4
    assert choose_storage(1000, 500) == 'factor' and choose_storage(1000, 501) == 'dense'
    assert choose_storage(3000, 750) == 'factor' and choose_storage(3000, 751) == 'dense'
    assert choose_storage(706, 706) == 'dense'


def test_c_abi_exports_every_declared_symbol():
    """libvilma_b200.so builds for sm_100a without a GPU and exports include/vilma_b200.h."""
    from vilma_b200 import _build, _lib
    _build.build_library()
    header = open(os.path.join(ROOT, 'include', 'vilma_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(vb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                      # getattr on every symbol
    assert lib.vb_abi_version() == 1
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r'\bT (vb_[a-z0-9_]+)', out))
    assert declared <= exported


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from vilma_b200._lib import VilmaB200Error
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    ld = BlockDiagonalMatrix([LowRankMatrix(X=np.eye(3))])
    with pytest.raises(VilmaB200Error):
        ld.dot(np.ones(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'vilma_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn


def test_constructor_validation_matches_reference_errors():
    """VIScheme.__init__ argument checks (reference variational_inference.py:143-203, :606-613)."""
    from _np_engine import NumpyShardEngine
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    fx = load_case('vischeme_unlinked_a1_s0_t0')
    lds = build_ld(fx, LowRankMatrix, BlockDiagonalMatrix)
    base = dict(vi_kwargs(fx), ld_mats=lds, precomputed=precomputed(fx),
                engine_factory=NumpyShardEngine)

    def build(**over):
        kw = dict(base)
        kw.update(over)
        return MultiPopVI(**kw)

    build()                                            # the unmodified arguments are fine
    for missing in ('init_hg', 'gwas_N', 'num_its', 'annotations', 'std_errs'):
        with pytest.raises(ValueError):
            build(**{missing: None})
    bad = fx['betas'].copy()
    bad[0, 3] = np.nan
    with pytest.raises(ValueError):
        build(marginal_effects=bad)
    bad = fx['std_errs'].copy()
    bad[1, 0] = np.inf
    with pytest.raises(ValueError):
        build(std_errs=bad)
    with pytest.raises(ValueError):
        build(ld_mats=lds[:1])                         # fewer LD matrices than populations
    with pytest.raises(ValueError):
        build(ld_mats=[lds[0], 'not an LD operator'])
    ann = np.array(fx['annotations'], dtype=float)
    ann[0, 0] = 0
    with pytest.raises(ValueError):
        build(annotations=ann)                         # a SNP without annotation
    with pytest.raises(ValueError):
        build(annotations=np.ones((7, 1)))             # wrong length
    with pytest.raises(ValueError):
        build(mixture_covs=[np.eye(3)])                # wrong shape
    with pytest.raises(ValueError):
        build(mixture_covs=[np.array([[1., 2.], [2., 1.]])])   # not positive definite


def test_update_beta_requires_nat_grad():
    """reference :770-772"""
    fx = load_case('vischeme_unlinked_a1_s0_t0')
    vi = make_host_vi(fx)
    params = (fx['init_vi_mu'], fx['init_vi_delta'], fx['init_hyper_delta'])
    with pytest.raises(RuntimeError):
        vi._update_beta(*params, None, np.ones(5), 0, 2.)


def test_checkpoint_files_written(tmp_path):
    """-checkpoint.<it>.npz cadence and contents (reference :362-367), host loop."""
    from _np_engine import NumpyShardEngine
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    from vilma_b200.variational_inference import MultiPopVI
    fx = load_case('vischeme_unlinked_a2_s0_t1')
    kw = dict(vi_kwargs(fx), checkpoint=True, checkpoint_freq=4, output=str(tmp_path / 'run'))
    vi = MultiPopVI(ld_mats=build_ld(fx, LowRankMatrix, BlockDiagonalMatrix),
                    precomputed=precomputed(fx), engine_factory=NumpyShardEngine, **kw)
    np.random.seed(int(fx['seed']))
    vi.optimize(None)
    its = len(fx['traj_L0'])
    expect = ['run-checkpoint.%d.npz' % i for i in range(0, its, 4)]
    assert sorted(os.listdir(tmp_path)) == sorted(expect)
    z = np.load(tmp_path / expect[0])
    assert sorted(z.files) == ['error_scaling', 'hyper_delta', 'scalings', 'vi_delta', 'vi_mu']
    assert np.allclose(z['vi_mu'], fx['init_vi_mu'], rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize('src', ['exp_check.c', 'log_check.c'])
def test_device_math_helpers_accuracy(src, tmp_path):
    """vb_exp_nonpos / vb_log_pos / vb_rcp_pos (csrc/vb_common.cuh) restated in C with the same
    operations (fma, Cody-Waite reduction, polynomial): <= 1 ulp from glibc over 2e7 arguments."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    exe = str(tmp_path / 'chk')
    subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-mfma', os.path.join(ROOT, 'tools', src),
                    '-o', exe, '-lm'], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert 'BAD' not in out
    import re
    for m in re.finditer(r'max ulp ([0-9.]+)', out):
        assert float(m.group(1)) <= 1.0, out


def test_rank_cpu_placement_plan():
    """dist.plan_affinity: every rank of a NUMA node gets its own physical cores (all their SMT
    threads), no overlap; too few cores -> leave the affinity alone."""
    from vilma_b200.dist import _parse_cpulist, plan_affinity
    assert _parse_cpulist('0-3,8,10-11') == [0, 1, 2, 3, 8, 10, 11]
    # 2 sockets x 8 cores x 2 threads: cpu c and c+16 are siblings; node 0 = cores 0-7
    topo = lambda c: (str((c % 16) // 8), str(c % 8))
    node0 = [c for c in range(32) if (c % 16) < 8]
    plans = [plan_affinity(range(32), node0, i, 4, topo) for i in range(4)]
    assert plans[0] == [0, 1, 16, 17] and plans[3] == [6, 7, 22, 23]
    flat = [c for p in plans for c in p]
    assert len(flat) == len(set(flat)) == 16
    # restricted cpuset: only what is allowed is used
    assert plan_affinity([0, 1, 2, 3], node0, 1, 2, topo) == [2, 3]
    # no overlap between node and cpuset -> fall back to the allowed set
    assert plan_affinity([8, 9, 24, 25], node0, 0, 2, topo) == [8, 24]
    assert plan_affinity(range(4), node0, 0, 8, topo) is None


def test_synth_lowrank_block_setup_values():
    """vilma_b200.synth (bench scaffolding for configs[2]/[4]): the low-rank block's factors, the
    set-up values computed from them and the Woodbury ridge start agree with dense NumPy algebra."""
    import torch
    from vilma_b200 import synth
    dev = torch.device('cpu')
    n = 96
    se = synth.block_se(n, 42, 3, 1e5, dev)
    beta = synth.shared_effects(n, 42, 3, 3, 1_200_000, dev)
    blk = synth.cohort_block(n, 1001, 3, 1, se, beta[1], dev, 0.3, 0.99)
    U, s = blk['U'].numpy(), blk['s'].numpy()
    R = (U * s) @ U.T
    assert blk['rank'] == int(np.ceil(0.3 * n)) - 1            # centred panel: one degree of freedom lost
    assert np.allclose(U.T @ U, np.eye(len(s)), atol=1e-10)
    assert np.allclose(np.diag(R), blk['ld_diag'].numpy(), rtol=1e-12)
    z = (blk['beta_hat'] / se).numpy()
    mle = np.linalg.pinv(R, rcond=1e-10) @ z
    assert np.isclose(blk['chi'], z @ mle, rtol=1e-9)
    assert np.allclose(blk['adj'].numpy(), (R @ mle) / se.numpy(), rtol=1e-8, atol=1e-8)
    prior = 3e-7
    ridge = synth.ridge_start(blk, se, prior).numpy()
    ref = np.linalg.solve(R + np.diag(se.numpy()**2 / prior), blk['adj'].numpy() * se.numpy()) * se.numpy()
    assert np.allclose(ridge, ref, rtol=1e-9, atol=1e-14)
    # full-rank dense block: the same quantities through the Cholesky route
    full = synth.cohort_block(n, 1000, 3, 0, se, beta[0], dev, 2.0, 1.0)
    assert full['U'] is None and full['rank'] == n
    Rf = full['R'].numpy()
    zf = (full['beta_hat'] / se).numpy()
    assert np.isclose(full['chi'], zf @ np.linalg.solve(Rf, zf), rtol=1e-9)


def test_bench_multi_cohort_grid():
    """bench.mixture_grid_multi: the reference's _make_simple with its indefinite members removed
    (the reference's own constructor rejects them), and the 256-matrix custom grid, are SPD."""
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(0)
    se = 10.0 ** rng.uniform(-3, -2, size=(3, 4000))
    beta = rng.normal(size=(3, 4000)) * se * 2
    covs = bench.mixture_grid_multi(beta, se, 3, ('simple', 2))
    assert 60 < len(covs) <= 123
    assert all(c.shape == (3, 3) and np.all(np.linalg.eigvalsh(c) > 0) for c in covs)
    se5, beta5 = np.tile(se[:1], (5, 1)), np.tile(beta[:1], (5, 1))
    covs = bench.mixture_grid_multi(beta5, se5, 5, ('custom', 256))
    assert len(covs) == 256
    assert all(np.all(np.linalg.eigvalsh(c) > 0) for c in covs)
    # seeded: the same grid every time
    again = bench.mixture_grid_multi(beta5, se5, 5, ('custom', 256))
    assert all(np.array_equal(a, b) for a, b in zip(covs, again))


def test_setup_thread_pool_matches_serial(monkeypatch):
    """vilma_b200._pool: per-block set-up work (eigh at load, pseudo-inverse, ridge solve, dense
    reconstruction for upload) on the thread pool gives what the serial order gives, in block order,
    and agrees with the oracle's operators."""
    from oracle.ld_np import BlockDiagonalLD, LowRankBlock
    from vilma_b200 import _pool
    from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
    monkeypatch.setattr(_pool, 'MIN_BLOCKS', 1)
    rng = np.random.default_rng(3)
    sizes = [40, 75, 33, 90, 61, 48, 57]
    mats = []
    for n in sizes:
        g = rng.normal(size=(2 * n, n))
        g -= g.mean(0)
        g /= np.sqrt((g * g).sum(0))
        mats.append(g.T @ g)
    M = sum(sizes) + 5
    perm = rng.permutation(M)
    missing = perm[sum(sizes):]

    def build(threads):
        monkeypatch.setenv('VILMA_B200_SETUP_THREADS', str(threads))
        pipe = _pool.OrderedPipeline(depth=2)
        for m in mats:
            pipe.submit(LowRankMatrix, m, 0.95)
        return BlockDiagonalMatrix(pipe.results(), perm=perm, missing=missing)

    serial, pooled = build(1), build(4)
    assert [m.shape for m in pooled.matrices] == [m.shape for m in serial.matrices]
    z = rng.normal(size=M)
    reg = rng.uniform(0.05, 0.2, size=M)
    monkeypatch.setenv('VILMA_B200_SETUP_THREADS', '4')
    inv_p, ridge_p = pooled.inverse.dot(z), pooled.ridge_inverse_dot(z, reg)
    blocks_p = pooled.device_blocks()
    monkeypatch.setenv('VILMA_B200_SETUP_THREADS', '1')
    inv_s, ridge_s = serial.inverse.dot(z), serial.ridge_inverse_dot(z, reg)
    blocks_s = serial.device_blocks()
    assert np.allclose(inv_p, inv_s, rtol=1e-10, atol=1e-12)
    assert np.allclose(ridge_p, ridge_s, rtol=1e-10, atol=1e-12)
    assert [b['kind'] for b in blocks_p] == [b['kind'] for b in blocks_s]
    for bp, bs in zip(blocks_p, blocks_s):
        key = 'R' if bp['kind'] == 'dense' else 'U'
        assert np.allclose(bp[key], bs[key], rtol=1e-10, atol=1e-12)
    # against the oracle's host operators
    ora = BlockDiagonalLD([LowRankBlock(X=m, t=0.95) for m in mats], perm=perm, missing=missing)
    assert np.allclose(inv_p, ora.inverse_dot(z) if hasattr(ora, 'inverse_dot') else ora.inverse.dot(z),
                       rtol=1e-8, atol=1e-10)
    assert np.allclose(ridge_p, ora.ridge_inverse_dot(z, reg), rtol=1e-8, atol=1e-10)
    # an exception inside a worker reaches the caller
    with pytest.raises(ValueError):
        monkeypatch.setenv('VILMA_B200_SETUP_THREADS', '4')
        _pool.map_blocks(lambda m: LowRankMatrix(m + np.triu(np.ones_like(m), 1), 1.0), mats)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the driver's CPU arm): exactly one line on stdout, the contract's
    keys, metric / unit identical to the GPU arm's."""
    import json
    env = dict(os.environ, BENCH_SAMPLE_BLOCKS='6')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '1', '--converge', '0'], capture_output=True,
                         text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.split('\n') if l.strip()]
    assert len(lines) == 1, res.stdout[:500]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'SNP-updates/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('CAVI SNP-updates/sec')
    for key in ('value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'data', 'config',
                'cpu_baseline', 'e2e'):
        assert key in d
    # the unmodified reference (oracle/_ref, numba) when it is installed here, else the NumPy port
    from oracle import ref_loader
    want = 'reference' if ref_loader.available()[0] else 'port'
    assert d['cpu_baseline']['kind'] == want and d['cpu_baseline']['cores'] >= 1
    if want == 'port':
        assert 'note' in d['cpu_baseline']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['value'] > 0


def test_tile_kernel_launch_plan():
    """tile_plan (csrc/vilma_b200.cu) through vb_debug_tile_plan: which per-SNP kernel runs for the
    shapes of BASELINE.json, with how many warps per 32-SNP tile, within the shared-memory limit."""
    import ctypes as C
    from vilma_b200 import _lib
    lib = _lib.load()

    def plan(P, K, M=1_200_000, akf=0, sms=148):
        W, grid, smem = C.c_int(), C.c_int(), C.c_int64()
        _lib.check(lib.vb_debug_tile_plan(P, K, M, akf, sms, C.byref(W), C.byref(grid), C.byref(smem)))
        return W.value, grid.value, smem.value

    assert plan(1, 14)[0] == 0 and plan(2, 8)[0] == 0          # single-cohort default grid: three-pass kernel
    assert plan(2, 582)[0] == 16                                # two cohorts at -K 12: 512 threads on 188 KB
    assert plan(3, 87)[0] == 4 and plan(3, 123)[0] == 4         # C3
    assert plan(5, 256)[0] == 4                                 # C5 (255 registers: 256 threads per SM)
    assert plan(6, 34)[0] == 1
    assert plan(6, 2000)[0] == 0      # 250 logit slots x 8 warps do not fit: thread-per-SNP online kernel
    for P, K in ((1, 40), (2, 42), (2, 582), (3, 123), (4, 31), (5, 256), (6, 34)):
        W, grid, smem = plan(P, K)
        assert W in (1, 2, 4, 8, 16) and 0 < smem <= 227 * 1024
        assert 0 < grid <= 148 * 16
    # a rank that owns few SNPs never launches more CTAs than tiles or partial rows
    W, grid, smem = plan(3, 87, M=100)
    assert grid == 1
    # forced modes
    _lib.check(lib.vb_set_option(b'snp_tile', 0))
    try:
        assert plan(3, 123)[0] == 0
        _lib.check(lib.vb_set_option(b'snp_tile', 8))
        assert plan(1, 14)[0] == 8
    finally:
        _lib.check(lib.vb_set_option(b'snp_tile', -1))


def test_setup_pool_releases_blas_limit_on_error(monkeypatch):
    """A failing block (bad schema, asymmetric matrix) must not leave BLAS pinned to one thread or worker
    threads behind."""
    from threadpoolctl import threadpool_info
    from vilma_b200 import _pool
    monkeypatch.setenv('VILMA_B200_SETUP_THREADS', '4')
    before = [(d.get('internal_api'), d.get('num_threads')) for d in threadpool_info()]

    def boom(x):
        if x == 3:
            raise ValueError('bad block')
        return x

    pipe = _pool.OrderedPipeline(depth=2)
    with pytest.raises(ValueError):
        try:
            for x in range(8):
                pipe.submit(boom, x)
            pipe.results()
        finally:
            pipe.close()
    assert [(d.get('internal_api'), d.get('num_threads')) for d in threadpool_info()] == before
    monkeypatch.setattr(_pool, 'MIN_BLOCKS', 1)
    with pytest.raises(ValueError):
        _pool.map_blocks(boom, range(8))
    assert [(d.get('internal_api'), d.get('num_threads')) for d in threadpool_info()] == before
    # and the happy path keeps order
    pipe = _pool.OrderedPipeline(depth=2)
    for x in range(20):
        pipe.submit(lambda v: v * v, x)
    assert pipe.results() == [v * v for v in range(20)]


def test_streamed_npz_equals_savez(tmp_path):
    """outputs.save_fit_npz: the final `.npz` written slice by slice -- vi_sigma
    never whole on the host -- is byte-compatible with the reference's np.savez call
    (vi_options.py:263-265): same members, shapes, dtypes and values under np.load."""
    from vilma_b200 import outputs
    fx = load_case('syn_p3')
    vi = make_host_vi(fx)
    vi.num_its = 3
    np.random.seed(int(fx['seed']))
    params = vi.optimize(None)
    K, P, M = vi.num_mix, vi.num_pops, vi.num_loci
    plan = outputs.slice_plan(K, P, M, slice_bytes=5 * 8 * P * P * M)
    assert plan[0] == (0, 5) and plan[-1][1] == K and all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    stats = {}
    path = outputs.save_fit_npz(str(tmp_path / 'streamed'), vi, params, slice_bytes=5 * 8 * P * P * M,
                                stats=stats)
    assert path.endswith('streamed.npz')
    assert stats['slices'] == len(plan) > 3
    assert stats['max_slice_bytes'] <= 5 * 8 * P * P * M          # never more than one slice per buffer
    ref = vi.create_dump_dict(params)
    ref['vi_sigma'] = vi.vi_sigma
    np.savez(str(tmp_path / 'whole'), **ref)
    a, b = np.load(path), np.load(str(tmp_path / 'whole.npz'))
    assert a.files == b.files == ['vi_mu', 'vi_delta', 'hyper_delta', 'error_scaling', 'scalings', 'vi_sigma']
    for k in b.files:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape
        assert np.array_equal(a[k], b[k]), k
    # a member larger than one slice (K = 1 per slice) and the degenerate single-slice plan
    assert outputs.slice_plan(7, 2, 10, slice_bytes=1) == [(k, k + 1) for k in range(7)]
    assert outputs.slice_plan(7, 2, 10, slice_bytes=1 << 40) == [(0, 7)]


def _partition_reference(block_lists, M, world):
    """The straightforward per-SNP formulation (union-find + one-by-one dealing) the vectorised
    partition_snps must reproduce exactly."""
    parent = list(range(M))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i
    for blocks in block_lists:
        for snps, _ in blocks:
            for i in snps[1:]:
                a, b = find(int(snps[0])), find(int(i))
                if a != b:
                    parent[max(a, b)] = min(a, b)
    roots = np.array([find(i) for i in range(M)])
    cost = np.zeros(M)
    in_block = np.zeros(M, dtype=bool)
    for blocks in block_lists:
        for snps, c in blocks:
            if len(snps):
                cost[roots[int(snps[0])]] += c
                in_block[snps] = True
    comp = np.unique(roots[in_block])
    order = comp[np.argsort(-cost[comp], kind='stable')]
    load = np.zeros(world)
    owner = np.full(M, -1)
    own_root = {}
    for r in order:
        k = int(np.argmin(load))
        own_root[int(r)] = k
        load[k] += cost[r]
    for i in np.where(in_block)[0]:
        owner[i] = own_root[int(roots[i])]
    counts = np.array([(owner == r).sum() for r in range(world)])
    for i in np.where(~in_block)[0]:
        k = int(np.argmin(counts))
        owner[i] = k
        counts[k] += 1
    return [np.where(owner == r)[0] for r in range(world)]


def test_partition_vectorised_matches_per_snp_formulation():
    from vilma_b200.partition import partition_snps
    rng = np.random.default_rng(3)
    for trial in range(6):
        M = int(rng.integers(200, 900))
        lists = []
        for p in range(int(rng.integers(1, 4))):
            # cohorts cut the (shuffled) SNPs at different places, and each leaves some SNPs out
            order = rng.permutation(M) if trial % 2 else np.arange(M)
            keep = order[rng.random(M) > 0.05]
            cuts = np.sort(rng.choice(np.arange(1, len(keep)), size=int(rng.integers(3, 12)), replace=False))
            blocks = [(b, float(len(b))**2) for b in np.split(keep, cuts)]
            lists.append(blocks)
        for world in (2, 3, 8):
            got = partition_snps(lists, M, world)
            want = _partition_reference(lists, M, world)
            assert all(np.array_equal(a, b) for a, b in zip(got, want)), (trial, world)


def test_partition_scales_to_benchmark_size():
    """1.2M SNPs in 1700 blocks (BASELINE configs[1]) partitions in well under the per-SNP
    interpreter loop's minutes."""
    import time
    from vilma_b200.partition import partition_snps
    from vilma_b200 import synth
    n = synth.block_sizes(1_188_000, 1700)
    starts = np.concatenate([[0], np.cumsum(n)])
    blocks = [(np.arange(starts[b], starts[b + 1]), float(n[b])**2) for b in range(len(n))]
    t0 = time.time()
    parts = partition_snps([blocks], 1_200_000, 8)
    assert time.time() - t0 < 10
    assert sum(len(p) for p in parts) == 1_200_000
    loads = [sum(float(n[b])**2 for b in range(len(n)) if starts[b] in set_) for set_ in
             [set(p[np.isin(p, starts[:-1])].tolist()) for p in parts]]
    assert max(loads) / min(loads) < 1.02


def test_index_runs_and_take_runs():
    """dist.take_runs (the host-side cut of a rank's shard) == ndarray.take for sorted indices."""
    from vilma_b200.dist import index_runs, take_runs
    idx = np.array([3, 4, 5, 9, 10, 20, 21, 22, 23, 40])
    assert index_runs(idx) == [(3, 6, 0), (9, 11, 3), (20, 24, 5), (40, 41, 9)]
    assert index_runs(np.array([], dtype=np.int64)) == []
    rng = np.random.default_rng(1)
    a = rng.standard_normal((4, 3, 50))
    assert np.array_equal(take_runs(a, idx, 2), a.take(idx, axis=2))
    b = rng.standard_normal((50, 7))
    assert np.array_equal(take_runs(b, idx, 0), b.take(idx, axis=0))
    scattered = np.arange(0, 50, 2)                    # no runs: falls back to a gather
    assert np.array_equal(take_runs(b, scattered, 0), b[scattered])
