"""Device engine: thin Python wrapper over the C ABI (include/vilma_b200.h).

PyTorch appears here only as the owner of small device/pinned buffers (the statistics
vectors that may be all-reduced over NCCL); all compute is in libvilma_b200.so.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


class DeviceContext:
    """One CUDA context handle (vb_ctx) per (process, device); shared by LD operators and fits."""

    _instances = {}

    def __init__(self, device):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.VilmaB200Error(
                'vilma_b200 needs a CUDA device (B200, sm_100a); none is visible and there is '
                'no CPU fallback.')
        self.lib = _lib.load()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        self.handle = C.c_void_p()
        self._lds = weakref.WeakSet()       # LD operators living in this context (closed before it)
        # stream 0 = the legacy default stream, which is also torch's default stream
        _lib.check(self.lib.vb_ctx_create(self.device, None, C.byref(self.handle)))

    def close(self):
        """Free everything the context owns on the device (fit state, mailbox, pinned staging) after
        closing the LD operators that were created in it.  Idempotent."""
        if self.handle:
            for ld in list(self._lds):
                ld.close()
            self.lib.vb_ctx_destroy(self.handle)
            self.handle = C.c_void_p()
            for dev, inst in list(DeviceContext._instances.items()):
                if inst is self:
                    del DeviceContext._instances[dev]

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def get(cls, device=None):
        if device is None:
            torch = _torch()
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def sync(self):
        _lib.check(self.lib.vb_ctx_sync(self.handle))

    def launch_count(self):
        return int(self.lib.vb_ctx_launch_count(self.handle))

    def profile(self, enable=True):
        _lib.check(self.lib.vb_ctx_profile(self.handle, 1 if enable else 0))

    def profile_read(self):
        """{'ld_matvec' | 'snp' | 'finish' | 'bookkeeping': (total_ms, launches)} since last reset."""
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        _lib.check(self.lib.vb_ctx_profile_read(self.handle, ms, cnt))
        return {'ld_matvec': (ms[0], cnt[0]), 'snp': (ms[1], cnt[1]), 'finish': (ms[2], cnt[2]),
                'bookkeeping': (ms[3], cnt[3])}


def sym_nmax():
    """Largest block stored symmetric-packed (csrc/ld_kernels.cuh VB_SYM_NMAX)."""
    return int(_lib.load().vb_ld_sym_nmax())


def dense_bytes(n):
    """Bytes one mat-vec streams for a dense block: symmetric-packed up to sym_nmax(), else full."""
    return 4 * n * (n + 1) if n <= sym_nmax() else 8 * n * n


def fac_nmax():
    """Largest factor block (rows) stored in the read-once form U sqrt(s) (0: switched off)."""
    return int(_lib.load().vb_ld_fac_nmax())


def factor_bytes(n, r):
    """Bytes one mat-vec streams for a factor block: 8 n r read once when the block fits the read-once
    kernel (n padded to even), else the two passes V' = diag(s) U^T and U."""
    return 8 * (n + (n & 1)) * r if n <= fac_nmax() else 16 * n * r


def choose_storage(n, r):
    """'dense' when streaming the (packed) n x n block costs no more than the factor form."""
    return 'dense' if dense_bytes(n) <= factor_bytes(n, r) else 'factor'


def set_option(name, value):
    _lib.check(_lib.load().vb_set_option(name.encode(), int(value)))


class DeviceLD:
    """One cohort's block-diagonal LD operator resident in HBM (vb_ld).

    blocks: list of dicts {'n', 'kind': 'dense'|'factor', 'R' | ('U','s')}; arrays may be
    host numpy (float64, C order) or torch CUDA tensors.  perm_local[j] = local SNP index of
    block-order position j.  M = number of SNPs on this rank.
    """

    def __init__(self, ctx, M, blocks=None, perm_local=None, n=None, rank=None):
        """Either pass `blocks` + `perm_local` (everything at once) or `n` / `rank` arrays and
        then call set_dense / set_factor per block and finalize(perm_local) (streaming build)."""
        self.ctx = ctx
        self.lib = ctx.lib
        self.M = int(M)
        if blocks is not None:
            n = np.array([b['n'] for b in blocks], dtype=np.int64)
            rank = np.array([-1 if b['kind'] == 'dense' else b['U'].shape[1] for b in blocks],
                            dtype=np.int64)
        n = np.ascontiguousarray(n, dtype=np.int64)
        rank = np.ascontiguousarray(rank, dtype=np.int64)
        self.handle = C.c_void_p()
        self.bytes = 0
        ctx._lds.add(self)
        _lib.check(self.lib.vb_ld_create(
            ctx.handle, self.M, len(n), n.ctypes.data_as(_lib.c_i64p),
            rank.ctypes.data_as(_lib.c_i64p), C.byref(self.handle)))
        if blocks is not None:
            for b, blk in enumerate(blocks):
                if blk['kind'] == 'dense':
                    self.set_dense(b, blk['R'])
                else:
                    self.set_factor(b, blk['U'], blk['s'])
            self.finalize(perm_local)

    def set_dense(self, b, R):
        ptr, on_dev, keep = self._ptr(R)
        _lib.check(self.lib.vb_ld_set_dense(self.handle, b, ptr, int(R.shape[1]), on_dev))

    def set_factor(self, b, U, s):
        pu, on_dev, keep = self._ptr(U)
        ps, on_dev2, keep2 = self._ptr(s)
        assert on_dev == on_dev2
        _lib.check(self.lib.vb_ld_set_factor(self.handle, b, pu, ps, on_dev))

    def finalize(self, perm_local):
        perm = np.ascontiguousarray(perm_local, dtype=np.int64)
        _lib.check(self.lib.vb_ld_finalize(self.handle, perm.ctypes.data_as(_lib.c_i64p),
                                           perm.shape[0]))
        self.bytes = int(self.lib.vb_ld_bytes(self.handle))

    @staticmethod
    def _ptr(a):
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, dtype=np.float64)
            return C.c_void_p(a.ctypes.data), 0, a
        torch = _torch()
        assert a.is_cuda and a.dtype == torch.float64
        a = a.contiguous()
        return C.c_void_p(a.data_ptr()), 1, a

    def dot(self, x):
        """y = R x for a host vector in this rank's SNP order."""
        torch = _torch()
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).cuda(self.ctx.device)
        yd = torch.empty_like(xd)
        _lib.check(self.lib.vb_ld_dot(self.handle, C.c_void_p(xd.data_ptr()),
                                      C.c_void_p(yd.data_ptr())))
        return yd.cpu().numpy()

    def dot_device(self, xd, yd):
        _lib.check(self.lib.vb_ld_dot(self.handle, C.c_void_p(xd.data_ptr()),
                                      C.c_void_p(yd.data_ptr())))

    def close(self):
        if self.handle:
            if self.ctx.handle:             # (a closed context has already closed its operators)
                self.lib.vb_ld_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaEngine:
    """Device-resident fit state + the fused kernels, for the SNPs owned by this rank.

    All arrays passed in are in this rank's SNP order.  Statistics come back as small torch
    CUDA tensors so the caller can all-reduce them before reading.
    """

    N_DIFF = 10

    def __init__(self, ctx, lds, K, P, M, A, adj, se, sld, scalings, annotations,
                 mixture_prec, log_det):
        torch = _torch()
        self.ctx, self.lib = ctx, ctx.lib
        self.lds = list(lds)
        self.K, self.P, self.M, self.A = int(K), int(P), int(M), int(A)
        arr = (C.c_void_p * P)(*[ld.handle for ld in self.lds])
        _lib.check(self.lib.vb_fit_create(ctx.handle, self.K, self.P, self.M, self.A,
                                          C.cast(arr, C.POINTER(C.c_void_p))))
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        adj, se, sld, scalings = f64(adj), f64(se), f64(sld), f64(scalings)
        ann = np.ascontiguousarray(annotations, dtype=np.int32)
        assert adj.shape == (P, M) and ann.shape == (M,)
        _lib.check(self.lib.vb_fit_set_snp_data(ctx.handle, _lib.np_ptr(adj), _lib.np_ptr(se),
                                                _lib.np_ptr(sld), _lib.np_ptr(scalings),
                                                _lib.np_ptr(ann)))
        prec, ld = f64(mixture_prec).reshape(K, P, P), f64(log_det)
        _lib.check(self.lib.vb_fit_set_mixture(ctx.handle, _lib.np_ptr(prec), _lib.np_ptr(ld)))
        dev = torch.device('cuda', ctx.device)
        self.n_stats = 3 * P + 3 + 48 + 10     # + fused annotation sums and convergence slots
        self._stats = torch.zeros(self.n_stats, dtype=torch.float64, device=dev)
        self._ann = torch.zeros(A * K, dtype=torch.float64, device=dev)
        self._diff = torch.zeros(self.N_DIFF, dtype=torch.float64, device=dev)
        self.ld_bytes = sum(ld_.bytes for ld_ in self.lds)
        self.native_ready = False

    # ---- small inputs
    def set_hyper(self, hyper):
        h = np.ascontiguousarray(hyper, dtype=np.float64)
        assert h.shape == (self.A, self.K)
        _lib.check(self.lib.vb_fit_set_hyper(self.ctx.handle, _lib.np_ptr(h)))

    def set_delta_grad(self, table):
        g = np.ascontiguousarray(table, dtype=np.float64)
        assert g.shape == (self.A, self.K - 1)
        if self.K > 1:
            _lib.check(self.lib.vb_fit_set_delta_grad(self.ctx.handle, _lib.np_ptr(g)))

    def set_tau(self, tau):
        t = np.ascontiguousarray(tau, dtype=np.float64)
        assert t.shape == (self.P,)
        _lib.check(self.lib.vb_fit_set_tau(self.ctx.handle, _lib.np_ptr(t)))

    # ---- state transfer (reference host layouts in, device layouts inside)
    def set_params(self, vi_mu, vi_delta):
        mu = np.ascontiguousarray(vi_mu, dtype=np.float64)
        dmk = np.ascontiguousarray(vi_delta, dtype=np.float64)
        assert mu.shape == (self.K, self.P, self.M) and dmk.shape == (self.M, self.K)
        _lib.check(self.lib.vb_fit_set_params(self.ctx.handle, _lib.np_ptr(mu), _lib.np_ptr(dmk)))

    @staticmethod
    def _host_array(shape):
        """A fresh float64 host array in page-locked memory (torch's caching host allocator makes
        repeated allocations cheap); device->host copies into it run at PCIe speed instead of the
        ~5 GB/s of a pageable destination."""
        torch = _torch()
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()

    def get_params(self, out_mu=None, out_delta=None):
        mu = self._host_array((self.K, self.P, self.M)) if out_mu is None else out_mu
        dmk = self._host_array((self.M, self.K)) if out_delta is None else out_delta
        _lib.check(self.lib.vb_fit_get_params(self.ctx.handle, _lib.np_ptr(mu), _lib.np_ptr(dmk)))
        return mu, dmk

    # ---- sharded transfers against GLOBAL host arrays the device can address (multi-GPU)
    def set_shard(self, snps, M_total):
        idx = np.ascontiguousarray(snps, dtype=np.int64)
        assert idx.shape == (self.M,)
        _lib.check(self.lib.vb_fit_set_shard(self.ctx.handle, _lib.np_ptr(idx), int(M_total)))
        self._shard_total = int(M_total)

    def host_accessible(self, arr):
        """True if kernels can read/write `arr` in place (page-locked or registered, C-contiguous float64)."""
        return (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags.c_contiguous
                and bool(self.lib.vb_host_accessible(C.c_void_p(arr.ctypes.data))))

    def host_register(self, address, nbytes):
        """Page-lock and map a host range (e.g. a node-shared mapping); False if the driver declines."""
        return self.lib.vb_host_register(C.c_void_p(address), int(nbytes)) == 0

    def host_unregister(self, address):
        return self.lib.vb_host_unregister(C.c_void_p(address)) == 0

    def set_params_shard(self, vi_mu_global, vi_delta_global):
        mu = np.ascontiguousarray(vi_mu_global, dtype=np.float64)
        dmk = np.ascontiguousarray(vi_delta_global, dtype=np.float64)
        assert mu.shape == (self.K, self.P, self._shard_total) and dmk.shape == (self._shard_total, self.K)
        _lib.check(self.lib.vb_fit_set_params_shard(self.ctx.handle, _lib.np_ptr(mu), _lib.np_ptr(dmk)))

    def get_params_shard(self, vi_mu_global, vi_delta_global):
        assert vi_mu_global.shape == (self.K, self.P, self._shard_total)
        assert vi_delta_global.shape == (self._shard_total, self.K)
        _lib.check(self.lib.vb_fit_get_params_shard(self.ctx.handle, C.c_void_p(vi_mu_global.ctypes.data),
                                                    C.c_void_p(vi_delta_global.ctypes.data)))

    def get_params_device(self):
        """(vi_mu [K,P,M], vi_delta [M,K]) of this rank as torch CUDA tensors."""
        torch = _torch()
        dev = torch.device('cuda', self.ctx.device)
        mu = torch.empty((self.K, self.P, self.M), dtype=torch.float64, device=dev)
        dmk = torch.empty((self.M, self.K), dtype=torch.float64, device=dev)
        _lib.check(self.lib.vb_fit_get_params_dev(self.ctx.handle, C.c_void_p(mu.data_ptr()),
                                                  C.c_void_p(dmk.data_ptr())))
        return mu, dmk

    def set_params_device(self, mu, dmk):
        assert mu.is_cuda and dmk.is_cuda and mu.is_contiguous() and dmk.is_contiguous()
        assert tuple(mu.shape) == (self.K, self.P, self.M) and tuple(dmk.shape) == (self.M, self.K)
        _lib.check(self.lib.vb_fit_set_params_dev(self.ctx.handle, C.c_void_p(mu.data_ptr()),
                                                  C.c_void_p(dmk.data_ptr())))

    # ---- evaluations (return the device stats tensor, valid until the next evaluation)
    def eval(self):
        _lib.check(self.lib.vb_fit_eval(self.ctx.handle, C.c_void_p(self._stats.data_ptr())))
        return self._stats

    def beta_trial(self, step):
        _lib.check(self.lib.vb_fit_beta_trial(self.ctx.handle, float(step),
                                              C.c_void_p(self._stats.data_ptr())))
        return self._stats

    def refresh_delta(self):
        _lib.check(self.lib.vb_fit_refresh_delta(self.ctx.handle,
                                                 C.c_void_p(self._stats.data_ptr())))
        return self._stats

    def accept(self):
        _lib.check(self.lib.vb_fit_accept(self.ctx.handle))

    def sum_annotations(self):
        _lib.check(self.lib.vb_fit_sum_annotations(self.ctx.handle,
                                                   C.c_void_p(self._ann.data_ptr())))
        return self._ann

    def init_delta(self, fake_mu):
        fm = np.ascontiguousarray(fake_mu, dtype=np.float64)
        assert fm.shape == (self.P, self.M)
        _lib.check(self.lib.vb_fit_init_delta(self.ctx.handle, _lib.np_ptr(fm)))

    def init_mu(self):
        _lib.check(self.lib.vb_fit_init_mu(self.ctx.handle))

    def posterior(self):
        pm = np.empty((self.P, self.M))
        pv = np.empty((self.P, self.M))
        _lib.check(self.lib.vb_fit_posterior(self.ctx.handle, _lib.np_ptr(pm), _lib.np_ptr(pv)))
        return pm, pv

    def pm_diff(self, atol, rtol):
        _lib.check(self.lib.vb_fit_pm_diff(self.ctx.handle, float(atol), float(rtol),
                                           C.c_void_p(self._diff.data_ptr())))
        return self._diff

    def pm_mark(self, which):
        _lib.check(self.lib.vb_fit_pm_mark(self.ctx.handle, int(which)))

    def vi_sigma(self, k0=0, k1=None, out=None):
        k1 = self.K if k1 is None else k1
        if out is None:
            out = np.empty((k1 - k0, self.P, self.P, self.M))
        assert out.shape == (k1 - k0, self.P, self.P, self.M) and out.flags['C_CONTIGUOUS']
        _lib.check(self.lib.vb_fit_vi_sigma(self.ctx.handle, k0, k1, _lib.np_ptr(out)))
        return out

    # ---- native control loop (one outer iteration per call)
    def set_constants(self, chi_stat, ld_ranks, annotation_counts, log_det, scale_se):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        c, r, n, ld = f64(chi_stat), f64(ld_ranks), f64(annotation_counts), f64(log_det)
        _lib.check(self.lib.vb_fit_set_constants(self.ctx.handle, _lib.np_ptr(c), _lib.np_ptr(r),
                                                 _lib.np_ptr(n), _lib.np_ptr(ld),
                                                 1 if scale_se else 0))
        self.native_ready = True

    def init_comm(self, comm):
        """Rank plumbing of the native loop: the library's own NCCL communicator over the ranks of
        `comm` plus (one node, <= 8 ranks) the NVLink mailbox rendezvous (vb_xr_*)."""
        import os
        if comm.world > 1:
            from .dist import pin_rank
            self.cpus = pin_rank(comm, self.ctx.device)
            buf = C.create_string_buffer(128)
            if comm.rank == 0:
                _lib.check(self.lib.vb_nccl_unique_id(buf))
            ident = comm.broadcast_bytes(buf.raw)
            _lib.check(self.lib.vb_comm_init(self.ctx.handle, comm.world, comm.rank, ident))
            # with several ranks every reduction is a rendezvous: let the annotation sums ride along
            _lib.check(self.lib.vb_fit_set_fusion(self.ctx.handle, 1))
        # (one rank: a separate annotation-sum pass per outer iteration -- 58 us on C2 -- measured
        # cheaper than carrying the sums in every evaluation, +27 us per per-SNP launch)
        if os.environ.get('VILMA_B200_NO_XRANK', '0') != '1' and comm.world <= 8:
            h = C.create_string_buffer(64)
            _lib.check(self.lib.vb_xr_create(self.ctx.handle, h))
            handles = b''.join(comm.allgather_bytes(h.raw))
            _lib.check(self.lib.vb_xr_open(self.ctx.handle, comm.world, comm.rank, handles))

    def iteration(self, io, tau, hyper, stats):
        """vb_fit_iteration: `io` is a _lib.StepIO; tau/hyper/stats are float64 arrays updated in place."""
        rc = self.lib.vb_fit_iteration(self.ctx.handle, C.byref(io), _lib.np_ptr(tau),
                                       _lib.np_ptr(hyper), _lib.np_ptr(stats))
        if rc == 2:
            raise RuntimeError('Encountered a numerical error.')
        _lib.check(rc)

    def close(self):
        """Free the fit state (parameter buffers, scratch) of this engine's context."""
        if self.ctx.handle:
            self.lib.vb_fit_destroy(self.ctx.handle)
