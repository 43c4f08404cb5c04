"""Latency of a tiny NCCL all-reduce (+ D2H + sync) -- what one rendezvous of the fit costs."""
import os
import time
import torch
import torch.distributed as dist

rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
x = torch.ones(16, dtype=torch.float64, device='cuda')
h = torch.empty(16, dtype=torch.float64).pin_memory()
for _ in range(50):
    dist.all_reduce(x)
torch.cuda.synchronize()
dist.barrier()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    dist.all_reduce(x)
    h.copy_(x, non_blocking=True)
    torch.cuda.current_stream().synchronize()
t1 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print('world %d  NCCL_ALGO=%s NCCL_PROTO=%s : allreduce+D2H+sync %.1f us/op (host clock), back-to-back allreduce %.1f us/op (device)' % (
        dist.get_world_size(), os.environ.get('NCCL_ALGO'), os.environ.get('NCCL_PROTO'),
        (t1 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3))
dist.destroy_process_group()
