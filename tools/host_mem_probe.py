#!/usr/bin/env python
"""Host-memory probe under torchrun: strided reads / copies of page-locked arrays before and after the
rank pins itself to its cores (diagnoses slow host passes on multi-rank runs)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(rank)
multi = 'RANK' in os.environ
if multi:
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))


def probe(tag, arrs):
    for name, a in arrs.items():
        flat = a.reshape(-1)
        t = time.perf_counter(); s = float(flat[::max(1, flat.size // 509)].sum()); t1 = time.perf_counter()
        b = np.empty_like(a); t2 = time.perf_counter(); np.copyto(b, a); t3 = time.perf_counter(); np.copyto(b, a); t4 = time.perf_counter()
        print('[rank %d] %-12s %-10s strided sum %.3f ms, copy to fresh pageable %.1f ms, again %.1f ms' % (
            rank, tag, name, (t1 - t) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)


n = 14 * 1200000
arrs = {'pinned': torch.zeros(n, dtype=torch.float64).pin_memory().numpy(), 'pageable': np.zeros(n)}
print('[rank %d] affinity %d cpus, numa_balancing=%s thp=%s' % (
    rank, len(os.sched_getaffinity(0)), open('/proc/sys/kernel/numa_balancing').read().strip()
    if os.path.exists('/proc/sys/kernel/numa_balancing') else '?',
    open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip()), flush=True)
probe('before pin', arrs)
probe('before pin 2', arrs)
if multi:
    from vilma_b200.dist import TorchComm, pin_rank
    cpus = pin_rank(TorchComm(), rank)
    print('[rank %d] pinned to %s' % (rank, cpus), flush=True)
    probe('after pin', arrs)
    probe('after pin 2', arrs)
    arrs2 = {'pinned-new': torch.zeros(n, dtype=torch.float64).pin_memory().numpy()}
    probe('new after pin', arrs2)
    dist.barrier()
    dist.destroy_process_group()
