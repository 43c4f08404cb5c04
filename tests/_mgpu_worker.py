"""Worker for tests/test_gpu_multi.py: one rank of a multi-GPU fit of a golden fixture (NCCL)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))

from _fixtures import build_ld, load_case, vi_kwargs  # noqa: E402
from vilma_b200.dist import TorchComm  # noqa: E402
from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix  # noqa: E402
from vilma_b200.variational_inference import MultiPopVI  # noqa: E402

for name in sys.argv[1:]:
    for native in (True, False):
        fx = load_case(name)
        vi = MultiPopVI(ld_mats=build_ld(fx, LowRankMatrix, BlockDiagonalMatrix), comm=TorchComm(),
                        device=rank, **vi_kwargs(fx))
        vi.use_native_loop = native
        assert 0 < len(vi._snps) < fx['betas'].shape[1]
        np.random.seed(int(fx['seed']))
        params = vi.optimize(None)
        tr = vi.trajectory
        assert tr['trials'] == fx['traj_trials'].tolist(), (name, native, tr['trials'])
        assert np.array_equal(np.array(tr['L0']), fx['traj_L0'])
        # (extra fixtures: floor relative to the trajectory's scale, as tests/test_gpu_extra.py)
        eatol = 1e-13 * np.abs(fx['traj_elbo_out']).max() if name.startswith('xtr_') else 0
        assert np.allclose(tr['elbo'], fx['traj_elbo_out'], rtol=1e-8, atol=eatol)
        assert np.allclose(params[0], fx['final_vi_mu'], rtol=1e-6,
                           atol=max(1e-9, 1e-13 * np.abs(fx['final_vi_mu']).max()))
        assert np.allclose(params[1], fx['final_vi_delta'], rtol=1e-6, atol=1e-12)
        assert np.allclose(params[2], fx['final_hyper_delta'], rtol=1e-6, atol=1e-12)
        assert np.allclose(vi.error_scaling, fx['final_error_scaling'], rtol=1e-8)
        assert np.allclose(vi.real_posterior_mean(*params), fx['final_post_mean'], rtol=1e-6, atol=1e-9)
        # the fitted parameters came back through the node-shared, page-locked mapping the GPUs scattered
        # their SNPs into (every rank sees all of them: compared above); uploading them again cuts this
        # rank's runs out again and must reproduce the state bit for bit
        final_elbo = vi.elbo(params)
        assert vi._eng.host_accessible(np.asarray(params[0])) and vi._eng.host_accessible(np.asarray(params[1]))
        vi._resident = None
        again = vi.elbo(params)          # (evaluated by the given-state kernel: same value, other rounding)
        assert np.isclose(again, final_elbo, rtol=1e-10, atol=0)
        back = vi._download()
        assert np.array_equal(back[0], params[0]) and np.array_equal(back[1], params[1])
        # a pageable, non-contiguous copy: same state
        vi._resident = None
        assert vi.elbo((np.asfortranarray(params[0]), np.array(params[1]), np.array(params[2]))) == again
        if rank == 0:
            print('ok', name, 'native' if native else 'python', flush=True)
dist.barrier()
dist.destroy_process_group()
