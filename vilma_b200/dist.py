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


def index_runs(idx):
    """Maximal runs of consecutive integers in the sorted index array `idx`:
    [(global start, global stop, local start), ...]."""
    idx = np.asarray(idx, dtype=np.int64)
    if len(idx) == 0:
        return []
    cuts = np.where(np.diff(idx) != 1)[0] + 1
    lo = np.concatenate([[0], cuts])
    hi = np.concatenate([cuts, [len(idx)]])
    return [(int(idx[a]), int(idx[b - 1]) + 1, int(a)) for a, b in zip(lo, hi)]


def take_runs(arr, idx, axis, alloc=None):
    """arr.take(idx, axis) for a sorted `idx`, as a few slice copies (whole LD blocks are contiguous
    SNP ranges) instead of a gather.  `alloc(shape)` supplies the output (e.g. page-locked memory)."""
    arr = np.asarray(arr)
    runs = index_runs(idx)
    shape = list(arr.shape)
    shape[axis] = len(idx)
    out = np.empty(shape, dtype=arr.dtype) if alloc is None else alloc(tuple(shape))
    if len(runs) > max(64, len(idx) // 16):          # scattered indices: a plain gather is as good
        np.take(arr, idx, axis=axis, out=out)
        return out
    sl = [slice(None)] * arr.ndim
    dst = [slice(None)] * arr.ndim
    for g0, g1, l0 in runs:
        sl[axis] = slice(g0, g1)
        dst[axis] = slice(l0, l0 + (g1 - g0))
        out[tuple(dst)] = arr[tuple(sl)]
    return out


def _to_numpy(t):
    if isinstance(t, np.ndarray):
        return np.array(t, dtype=np.float64)
    return t.detach().cpu().numpy().astype(np.float64, copy=True)


_SHM_POOL = []      # node-shared result mappings of this process (TorchComm.shared_arrays)


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
        # (C-contiguous in the caller's axis order: a moved-axis view makes every later pass a transposing copy)
        return np.ascontiguousarray(np.moveaxis(out.reshape((M,) + rest), 0, axis))

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

    def shared_arrays(self, shapes):
        """float64 arrays of the given shapes inside ONE node-shared host mapping (a file in /dev/shm
        mapped by every rank): what one rank writes, every rank reads -- each rank fills in its own shard
        and all of them end up with the same zero-copy view, instead of an all-gather to every GPU
        followed by N full device->host copies.  Mappings are pooled: one whose arrays have been dropped
        on EVERY rank is handed out again (its pages are already faulted in and, where the caller had it
        page-locked, still registered); when none of the right size is free, the unreferenced ones are
        released and a new one is created.  Returns (arrays, entry) -- `entry`
        is the pool record, a dict the caller may keep notes in ('address', 'nbytes' are set) -- or None
        when the ranks do not share a host or /dev/shm lacks the space.  Collective."""
        import mmap
        import os
        import socket
        import weakref
        import torch
        shapes = [tuple(int(v) for v in shp) for shp in shapes]
        offs, total = [], 0
        for shp in shapes:
            offs.append(total)
            total += int(np.prod(shp)) * 8
            total = (total + 4095) & ~4095
        if not hasattr(self, '_shm_same_host'):
            # (VILMA_B200_NO_SHM=1 on any rank switches the shared mapping off for all of them)
            hosts = self.allgather_bytes((socket.gethostname() + '|' + os.environ.get('VILMA_B200_NO_SHM', '0')).encode())
            self._shm_same_host = len(set(hosts)) == 1 and hosts[0].endswith(b'|0')
        if not self._shm_same_host:
            return None
        # The pool belongs to the process, not to this communicator object.  A caller may have page-locked a
        # mapping with the CUDA driver (entry['release'] undoes that): a mapping is only unmapped through
        # its release hook -- unmapping a registered range would leave a stale registration behind that a
        # later mapping at the same address inherits.
        pool = self._shm_pool = _SHM_POOL
        # mappings that nobody references any more -- on any rank
        flag = torch.zeros(len(pool) + 1, dtype=torch.int32)
        for i, e in enumerate(pool):
            if all(r() is None for r in e['refs']):
                flag[i] = 1
        if self._backend == 'nccl':
            flag = flag.cuda()
        self._dist.all_reduce(flag, op=self._dist.ReduceOp.MIN)
        flag = flag.cpu().numpy()
        free = [e for i, e in enumerate(pool) if flag[i] == 1]
        entry = next((e for e in free if e['nbytes'] == total), None)
        if entry is None:
            # nothing of this size to hand out again: give the unreferenced mappings back first
            # (every rank sees the same flags, so the pools stay identical)
            for e in free:
                try:
                    if e.get('release') is not None:
                        e['release']()
                    e['mm'].close()
                except (BufferError, ValueError, OSError):
                    pass                          # (cannot unmap: it is dropped from the pool all the same --
                pool.remove(e)                    #  every rank's pool must keep the same entries)
        if entry is None:
            name = b''
            if self.rank == 0:
                try:
                    st = os.statvfs('/dev/shm')
                    if st.f_bavail * st.f_frsize > total + (64 << 20):
                        self._shm_seq = getattr(self, '_shm_seq', 0) + 1
                        name = ('/dev/shm/vilma_b200_%d_%d' % (os.getpid(), self._shm_seq)).encode()
                        fd = os.open(name.decode(), os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                        os.ftruncate(fd, total)
                        os.close(fd)
                except OSError:
                    name = b''
            name = self.broadcast_bytes(name)
            if not name:
                return None
            fd = os.open(name.decode(), os.O_RDWR)
            try:
                mm = mmap.mmap(fd, total)
            finally:
                os.close(fd)
            self.barrier()                       # everybody has it mapped
            if self.rank == 0:
                os.unlink(name.decode())         # the mappings live on
            base = np.frombuffer(mm, dtype=np.uint8)
            entry = {'mm': mm, 'nbytes': total, 'refs': [], 'address': int(base.ctypes.data)}
            del base
            self._shm_pool.append(entry)
        outs = [np.frombuffer(entry['mm'], dtype=np.float64, count=int(np.prod(shp)), offset=off).reshape(shp)
                for shp, off in zip(shapes, offs)]
        entry['refs'] = [weakref.ref(o.base if o.base is not None else o) for o in outs]
        return outs, entry

    @staticmethod
    def fill_shared(globals_, locals_, snps, axes):
        """Host copy of this rank's shards into the shared arrays (runs of consecutive SNPs)."""
        runs = index_runs(snps)
        for g, a, ax in zip(globals_, locals_, axes):
            src = [slice(None)] * a.ndim
            dst = [slice(None)] * a.ndim
            for g0, g1, l0 in runs:
                dst[ax] = slice(g0, g1)
                src[ax] = slice(l0, l0 + (g1 - g0))
                g[tuple(dst)] = a[tuple(src)]

    def gather_shared(self, locals_, snps, M, axes):
        """shared_arrays + fill_shared + barrier: global arrays assembled from per-rank host shards."""
        shapes = []
        for a, ax in zip(locals_, axes):
            shp = list(a.shape)
            shp[ax] = M
            shapes.append(shp)
        got = self.shared_arrays(shapes)
        if got is None:
            return None
        outs, _ = got
        self.fill_shared(outs, locals_, snps, axes)
        self.barrier()                       # every shard is in place
        return outs

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


# ---------------------------------------------------------------------------------------------
# CPU placement.  Every rank's launch thread spins on a mapped flag between evaluations; two ranks
# sharing a physical core (SMT siblings) or migrating between sockets were measured to make the
# whole job wait for them (8 GPUs: two of eight ranks with 40 % slower launches and 0.2 ms of extra
# host time per outer iteration).  Each rank therefore takes its own physical cores, on the NUMA node
# of its GPU when the kernel reports one.
# ---------------------------------------------------------------------------------------------
def _read(path):
    try:
        with open(path) as fh:
            return fh.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = []
    for part in (text or '').split(','):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(pci_bus_id):
    """NUMA node of a PCI device ('0000:1b:00.0'), or -1 when unknown (VMs, containers)."""
    val = _read('/sys/bus/pci/devices/%s/numa_node' % pci_bus_id.lower())
    try:
        return int(val)
    except (TypeError, ValueError):
        return -1


def physical_cores(cpus, topology=None):
    """Group logical CPUs into physical cores: list of lists, in order of first appearance.
    topology(cpu) -> (package, core) defaults to /sys/devices/system/cpu/cpuN/topology."""
    if topology is None:
        def topology(c):
            base = '/sys/devices/system/cpu/cpu%d/topology/' % c
            return (_read(base + 'physical_package_id'), _read(base + 'core_id') or str(c))
    cores = {}
    for c in cpus:
        cores.setdefault(topology(c), []).append(c)
    return list(cores.values())


def plan_affinity(allowed, node_cpus, index, count, topology=None):
    """CPUs for the `index`-th of `count` ranks that share a NUMA node: an equal share of the
    node's physical cores (all their hardware threads).  None = leave the affinity alone (fewer
    physical cores than ranks)."""
    cpus = [c for c in node_cpus if c in set(allowed)] or list(allowed)
    cores = physical_cores(sorted(cpus), topology)
    per = len(cores) // max(count, 1)
    if per < 1:
        return None
    mine = cores[index * per:(index + 1) * per]
    return sorted(c for core in mine for c in core)


def pin_rank(comm, device):
    """Pin this process (all its current threads; later ones inherit) to its share of cores.
    Returns the CPU list or None.  VILMA_B200_NO_PIN=1 disables."""
    import os
    if comm.world <= 1 or os.environ.get('VILMA_B200_NO_PIN', '0') == '1' \
            or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        import socket
        import torch
        pr = torch.cuda.get_device_properties(device)
        bus = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = gpu_numa_node(bus)
        # ranks of this host that share my node, in rank order
        peers = comm.allgather_bytes(repr((socket.gethostname(), node)).encode())
        same = [r for r, p in enumerate(peers) if p == peers[comm.rank]]
        allowed = sorted(os.sched_getaffinity(0))
        node_cpus = _parse_cpulist(_read('/sys/devices/system/node/node%d/cpulist' % node)) \
            if node >= 0 else allowed
        cpus = plan_affinity(allowed, node_cpus, same.index(comm.rank), len(same))
        if not cpus:
            return None
        for tid in os.listdir('/proc/self/task'):
            try:
                os.sched_setaffinity(int(tid), cpus)
            except OSError:
                pass
        return cpus
    except Exception:
        return None


def default_comm():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except Exception:
        pass
    return SingleComm()
