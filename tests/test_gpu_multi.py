"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): LD blocks sharded over ranks, the
statistics summed over NVLink (mailbox exchange in the native loop, NCCL in the Python loop) -- the
fit must reproduce the single-process reference trajectory on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# xtr_p3_ann48: three cohorts with A*K = 48 fused annotation sums -> 70 values per exchanged statistics
# vector, the mailbox at its largest (an earlier build sized its rows for 64); 16 LD blocks, so that
# eight ranks all own LD.  syn_p1_dense has exactly 8 blocks.
CASES = {2: ['syn_p1_dense', 'syn_p2_lowrank', 'syn_p5', 'xtr_p3_ann48'],
         4: ['syn_p1_dense', 'syn_p2_lowrank', 'syn_p5', 'xtr_p3_ann48'],
         8: ['syn_p1_dense', 'xtr_p3_ann48']}


@pytest.mark.parametrize('options', ['', 'snp_tile=2'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_fit_matches_reference(world, options):
    if _ngpus() < world:
        pytest.skip('needs %d GPUs' % world)
    env = dict(os.environ, VILMA_B200_OPTIONS=options)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29700 + world),
           os.path.join(ROOT, 'tests', '_mgpu_worker.py')] + CASES[world]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count('ok ') == 2 * len(CASES[world])
