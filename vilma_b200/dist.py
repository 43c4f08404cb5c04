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
        parts = [None] * self.world
        self._dist.all_gather_object(parts, (np.asarray(snps), np.asarray(local)))
        shape = list(local.shape)
        shape[axis] = M
        out = np.zeros(shape, dtype=local.dtype)
        for idx, arr in parts:
            sl = [slice(None)] * out.ndim
            sl[axis] = idx
            out[tuple(sl)] = arr
        return out

    def barrier(self):
        self._dist.barrier()

    def broadcast_bytes(self, data):
        box = [data]
        self._dist.broadcast_object_list(box, src=0)
        return box[0]


def default_comm():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except Exception:
        pass
    return SingleComm()
