"""Rank plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo in CPU tests).

The fit shards whole LD blocks (and the SNPs they cover) across ranks; the only data-path
exchange is an all-reduce of the small per-state statistics vector, after which every rank
takes the same host-side decision from identical reduced values.
"""
import numpy as np


class SingleComm:
    rank = 0
    world = 1

    def sum(self, t):
        """t: torch tensor or numpy array of local sums -> host numpy array of global sums."""
        return _to_numpy(t)

    def max(self, t):
        return _to_numpy(t)

    def gather_snp_axis(self, local, snps, M, axis):
        """Assemble a global array from per-rank shards along the SNP axis."""
        return local

    def barrier(self):
        pass

    def broadcast_bytes(self, data):
        return data

    def allgather_bytes(self, data):
        return [data]


def _to_numpy(t):
    if isinstance(t, np.ndarray):
        return np.array(t, dtype=np.float64)
    return t.detach().cpu().numpy().astype(np.float64, copy=True)


class TorchComm(SingleComm):
    """All-reduce through the default torch.distributed process group."""

    def __init__(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self._dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self._backend = dist.get_backend()

    def _reduce(self, t, op):
        import torch
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.array(t, dtype=np.float64))
            if self._backend == 'nccl':
                t = t.cuda()
        elif self._backend == 'gloo' and t.is_cuda:
            t = t.cpu()
        else:
            t = t.clone()
        self._dist.all_reduce(t, op=op)
        return t.cpu().numpy()

    def sum(self, t):
        return self._reduce(t, self._dist.ReduceOp.SUM)

    def max(self, t):
        return self._reduce(t, self._dist.ReduceOp.MAX)

    def gather_snp_axis(self, local, snps, M, axis):
        """All-gather per-rank shards (padded to a common length) as tensors -- device tensors
        over NCCL, host tensors over gloo -- and scatter them into the global array."""
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if self._backend == 'nccl' \
            else torch.device('cpu')
        local = np.moveaxis(np.asarray(local), axis, 0)
        rest = local.shape[1:]
        flat = np.ascontiguousarray(local.reshape(local.shape[0], -1))
        n_loc = torch.tensor([flat.shape[0]], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n_loc) for _ in range(self.world)]
        self._dist.all_gather(sizes, n_loc)
        sizes = [int(t.item()) for t in sizes]
        n_max = max(sizes)
        pad = torch.zeros((n_max, flat.shape[1]), dtype=torch.from_numpy(flat).dtype, device=dev)
        pad[:flat.shape[0]] = torch.from_numpy(flat).to(dev)
        idx = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
        idx[:flat.shape[0]] = torch.from_numpy(np.asarray(snps, dtype=np.int64)).to(dev)
        all_pad = [torch.empty_like(pad) for _ in range(self.world)]
        all_idx = [torch.empty_like(idx) for _ in range(self.world)]
        self._dist.all_gather(all_pad, pad)
        self._dist.all_gather(all_idx, idx)
        # scatter on the gathering device (one index_copy_), then a single copy to the host
        rows = torch.cat([all_pad[r][:sizes[r]] for r in range(self.world)], dim=0)
        where = torch.cat([all_idx[r][:sizes[r]] for r in range(self.world)], dim=0)
        full = torch.zeros((M, flat.shape[1]), dtype=rows.dtype, device=dev)
        full.index_copy_(0, where, rows)
        out = full.cpu().numpy()
        return np.moveaxis(out.reshape((M,) + rest), 0, axis)

    def gather_device(self, local, snps, M, axis):
        """gather_snp_axis for a torch tensor that already lives on the gathering device; the
        result is returned as a page-locked host array (one device->host copy)."""
        import torch
        local = local.movedim(axis, 0).contiguous()
        rest = tuple(local.shape[1:])
        flat = local.reshape(local.shape[0], -1)
        dev = flat.device
        n_loc = torch.tensor([flat.shape[0]], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n_loc) for _ in range(self.world)]
        self._dist.all_gather(sizes, n_loc)
        sizes = [int(t.item()) for t in sizes]
        n_max = max(sizes)
        pad = torch.zeros((n_max, flat.shape[1]), dtype=flat.dtype, device=dev)
        pad[:flat.shape[0]] = flat
        idx = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
        idx[:flat.shape[0]] = torch.as_tensor(np.asarray(snps, dtype=np.int64), device=dev)
        all_pad = [torch.empty_like(pad) for _ in range(self.world)]
        all_idx = [torch.empty_like(idx) for _ in range(self.world)]
        self._dist.all_gather(all_pad, pad)
        self._dist.all_gather(all_idx, idx)
        rows = torch.cat([all_pad[r][:sizes[r]] for r in range(self.world)], dim=0)
        where = torch.cat([all_idx[r][:sizes[r]] for r in range(self.world)], dim=0)
        full = torch.zeros((M, flat.shape[1]), dtype=rows.dtype, device=dev)
        full.index_copy_(0, where, rows)
        full = full.reshape((M,) + rest).movedim(0, axis).contiguous()
        host = torch.empty(full.shape, dtype=full.dtype, pin_memory=dev.type == 'cuda')
        host.copy_(full)
        return host.numpy()

    def barrier(self):
        self._dist.barrier()

    def broadcast_bytes(self, data):
        box = [data]
        self._dist.broadcast_object_list(box, src=0)
        return box[0]

    def allgather_bytes(self, data):
        out = [None] * self.world
        self._dist.all_gather_object(out, data)
        return out


def default_comm():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except Exception:
        pass
    return SingleComm()
