"""Block-diagonal LD operators: host mirror of ``vilma.matrix_structures``.

Same classes, fields and method names as the reference
(/root/reference/src/vilma/matrix_structures.py) so callers and tests read the same:

* ``LowRankMatrix``      one block held as ``U diag(s) V + diag(D)``          (:38-234)
* ``BlockDiagonalMatrix`` blocks + ``perm`` / ``inv_perm`` / ``missing``      (:237-447)

What runs where:

* ``BlockDiagonalMatrix.dot`` -- the hot operator (:389-408 -> :148-152) -- runs on the
  GPU through ``vb_ld_dot`` (TMA-staged fp64 mat-vec, ``csrc/ld_kernels.cuh``).  No CPU path.
* ``inverse.dot``, ``ridge_inverse_dot``, ``diag``, ``get_rank`` are one-off setup calls of
  ``VIScheme.__init__`` (variational_inference.py:189-252); they stay on host NumPy/LAPACK
  so their results match the reference's LAPACK path (SURVEY.md section 8f ranks moving them
  to the GPU as "next").
"""
import logging

import numpy as np


def _svd_threshold(matrix, ld_thresh):
    """Eigen-decompose and keep eigenvalues >= 1 - sqrt(t)   (matrix_structures.py:15-28)."""
    vals, vecs = np.linalg.eigh(matrix)
    keep = np.where(vals >= 1 - np.sqrt(ld_thresh))[0]
    if len(keep) == 0:
        n = matrix.shape[0]
        return np.ones((n, 1)), np.zeros(1), np.ones((1, n))
    u_mat = np.copy(vecs[:, keep])
    return u_mat, np.copy(vals[keep]), np.copy(u_mat.T)


class LowRankMatrix():
    """Low rank plus diagonal representation of a symmetric block.

    `lazy=True` (only with a matrix X and t >= 1, i.e. no truncation) postpones the
    eigendecomposition until something reads `u`, `s`, `v` or `inv_s`: the GPU set-up
    (BlockDiagonalMatrix.device_setup) certifies per block that no eigenvalue would be dropped and
    then never needs the factors -- the block's operator is X itself."""

    _LAZY = ('u', 's', 'v', 'inv_s')

    def __init__(self, X=None, t=1.0, u=None, s=None, v=None, D=None, hdf_file=None, lazy=False):
        self._X = None
        self.full_rank_certified = False       # set by the GPU set-up: every eigenvalue is kept
        if X is not None and lazy and t >= 1.0:
            if any(a is not None for a in (u, s, v, D)):
                raise ValueError('Cannot provide both a matrix and an SVD decomposition')
            X = np.asarray(X, dtype=np.float64)
            if not np.allclose(X, X.T):
                raise ValueError('Provided matrix is not symmetric')
            if hdf_file is not None:
                logging.info('hdf_file is ignored: vilma_b200 keeps LD factors in HBM, not on disk')
            self._X = X
            self._t = t
            self.D = np.zeros(X.shape[0])
            self.shape = X.shape
            return
        if X is not None:
            if any(a is not None for a in (u, s, v, D)):
                raise ValueError('Cannot provide both a matrix and an SVD decomposition')
            if not np.allclose(X, X.T):
                raise ValueError('Provided matrix is not symmetric')
            u, s, v = _svd_threshold(X, t)
            D = np.zeros(X.shape[0])
        else:
            if any(a is None for a in (u, s, v, D)):
                raise ValueError('Need to provide either a matrix or an SVD decomposition')
            sel = np.where(s >= 1 - np.sqrt(t))[0]
            u, s, v = u[:, sel], s[sel], v[sel, :]
        if hdf_file is not None:
            logging.info('hdf_file is ignored: vilma_b200 keeps LD factors in HBM, not on disk')
        self.D = np.array(D, dtype=np.float64)
        self._set_factors(u, s, v)

    def __getattr__(self, name):
        # only reached when the attribute is missing: the postponed factors of a lazy block
        if name in LowRankMatrix._LAZY and self.__dict__.get('_X') is not None:
            u, s, v = _svd_threshold(self._X, self._t)
            self._set_factors(u, s, v)
            return self.__dict__[name]
        raise AttributeError(name)

    @property
    def factorized(self):
        return 'u' in self.__dict__

    def symmetric_matrix(self):
        """The block as LAPACK's eigh sees it in the reference (lower triangle mirrored), or None when
        the block was not given as a matrix."""
        if self._X is None:
            return None
        low = np.tril(self._X)
        return low + np.tril(self._X, -1).T

    def _set_factors(self, u, s, v):
        big = s > (1e-12 * np.max(s))
        if big.sum() > 0:
            self.u = np.array(u[:, big], dtype=np.float64)
            self.s = np.array(s[big], dtype=np.float64)
            self.v = np.array(v[big, :], dtype=np.float64)
            self.inv_s = 1 / self.s
        else:
            self.u = np.array(u[:, 0, None], dtype=np.float64)
            self.s = np.zeros(1)
            self.v = np.array(v[None, 0, :], dtype=np.float64)
            self.inv_s = np.zeros(1)
        self.shape = (self.u.shape[0], self.v.shape[1])

    def _host_dot(self, vector):
        """U (s * (V x)) + D x on the host.  Setup-only helper (ridge / pseudo-inverse)."""
        mid = (self.s * self.v.dot(vector).T).T
        return self.u.dot(mid) + (self.D * vector.T).T

    def dot(self, vector):
        """Matrix @ vector on the GPU (single-block operator)."""
        return BlockDiagonalMatrix([self]).dot(vector)

    def inverse_dot(self, vector):
        """PseudoInverse(Matrix) @ vector   (matrix_structures.py:159-196), host LAPACK."""
        if np.any(np.isclose(np.abs(self.D), 0)):
            if np.all(np.isclose(self.D, 0)):
                return self.v.T.dot(self.u.T.dot(vector) * self.inv_s)
            full = np.diag(self.D) + (self.u * self.s).dot(self.v)
            e_vals = np.linalg.eigh(full)[0][::-1]
            hit = np.where(np.isclose(np.cumsum(e_vals) / np.sum(e_vals), 1.))[0]
            pos = hit[0] if len(hit) > 0 else len(e_vals) - 1
            rcond = e_vals[pos] / e_vals[0] * 0.1
            return np.linalg.pinv(full, rcond=rcond).dot(vector)
        small = np.diag(self.inv_s) + self.v.dot((self.u.T / self.D).T)
        small = np.linalg.inv(small)
        out = self.u.dot(small.dot(self.v.dot(vector / self.D)))
        out /= self.D
        return vector / self.D - out

    def diag(self):
        if self._X is not None and not self.factorized:
            return np.diag(self._X).astype(np.float64) + self.D      # every eigenpair kept: the matrix itself
        return np.einsum('ik,ki->i', self.u * self.s, self.v) + self.D

    def matrix_power(self, power):
        if not np.allclose(self.D, 0):
            raise NotImplementedError('Matrix powers with a non-zero diagonal part are not '
                                      'implemented.')
        return LowRankMatrix(u=self.u, s=self.s**power, v=self.v, D=self.D)

    def get_rank(self):
        if self.full_rank_certified and not self.factorized:
            return self.shape[0]
        if np.allclose(self.D, 0):
            if self.s.shape[0] > 1:
                return self.s.shape[0]
            return 0 if self.s[0] == 0 else 1
        if np.all(self.D > 0):
            return self.D.shape[0]
        full = np.diag(self.D) + np.einsum('ik,k,kj->ij', self.u, self.s, self.v)
        return np.linalg.matrix_rank(full, hermitian=True)


class BlockDiagonalMatrix():
    """Symmetric block-diagonal matrix whose blocks are LowRankMatrix objects."""

    def __init__(self, matrices, inverse=False, perm=None, missing=None):
        if missing is None:
            missing = np.array([], dtype=np.int64)
        self.missing = np.array(missing, dtype=np.int64)
        for matrix in matrices:
            if not isinstance(matrix, LowRankMatrix):
                raise ValueError('Component matrices must be of type LowRankMatrix')
        self.matrices = matrices
        self._inverted = inverse
        self.starts = np.cumsum([0] + [m.shape[0] for m in matrices])
        n = int(self.starts[-1]) + self.missing.shape[0]
        self.shape = (n, n)
        if perm is None:
            self.perm = np.arange(n)
        else:
            perm = np.asarray(perm)
            if perm.shape[0] != n:
                raise ValueError('perm must be a vector conformal to the non-missing parts '
                                 'of the matrix.')
            self.perm = np.array(perm, dtype=np.int64)
        self.inv_perm = np.argsort(self.perm)
        if not np.allclose(self.perm[self.inv_perm], np.arange(n)):
            raise ValueError('perm and missing should together contain all of the indices. '
                             'Some are missing.')
        self._device = {}      # id(ctx) -> DeviceLD

    # ---- device side -------------------------------------------------------------
    def device_blocks(self, block_ids=None):
        """Describe blocks for upload: dense reconstruction when rank is close to n, else factor.
        The reconstructions (one GEMM per block) run on the set-up thread pool."""
        from ._pool import map_blocks
        from .engine import choose_storage, fac_nmax
        nmax = fac_nmax()   # (loads the library on this thread before the workers ask for it)
        ids = range(len(self.matrices)) if block_ids is None else block_ids
        return map_blocks(lambda b: self._device_block(self.matrices[b], choose_storage, nmax), ids)

    @staticmethod
    def _device_block(m, choose_storage, fac_nmax=0):
        if not np.all(m.D == 0):
            raise NotImplementedError('device LD blocks must have a zero diagonal part D')
        if m.full_rank_certified and not m.factorized:
            # the GPU set-up showed that the reference keeps every eigenpair: U diag(s) U^T is X
            return {'n': m.shape[0], 'kind': 'dense', 'R': m.symmetric_matrix()}
        n, r = m.u.shape
        if m.s.shape[0] == 1 and m.s[0] == 0:
            # rank-0 dummy block (matrix_structures.py:141-145): the zero matrix
            return {'n': n, 'kind': 'factor', 'U': np.zeros((n, 1)), 's': np.zeros(1)}
        if choose_storage(n, r) == 'dense':
            return {'n': n, 'kind': 'dense', 'R': (m.u * m.s).dot(m.v)}
        # the reference multiplies by v (= u^T for every block built from X); keep its semantics
        # exactly for a caller-supplied v by folding v into the factor only when it is the
        # transpose, else fall back to the dense product.
        # (the read-once factor form stores U sqrt(s): a caller-made block with a negative weight goes dense)
        if m.v.shape == m.u.T.shape and np.array_equal(m.v, m.u.T) and (n > fac_nmax or m.s.min() >= 0):
            return {'n': n, 'kind': 'factor', 'U': m.u, 's': m.s}
        return {'n': n, 'kind': 'dense', 'R': (m.u * m.s).dot(m.v)}

    def to_device(self, ctx=None):
        """Upload (once per context) and return the DeviceLD of the whole operator."""
        from .engine import DeviceContext, DeviceLD
        if ctx is None:
            ctx = DeviceContext.get()
        key = id(ctx)
        if key not in self._device:
            nreal = int(self.starts[-1])
            self._device[key] = DeviceLD(ctx, self.shape[0], self.device_blocks(),
                                         self.perm[:nreal])
        return self._device[key]

    def device_setup(self, z_scores, regularizer, ctx=None):
        """Set-up products of VIScheme.__init__ (variational_inference.py:236-252) for this operator:

            mle   = pinv(R) z            (self.inverse.dot)
            rmle  = R mle                (self.dot)
            ridge = (R + diag(reg))^-1 rmle   (self.ridge_inverse_dot)
            chi   = z . mle,   rank = rank(R)

        Dense blocks that were loaded lazily go through the GPU (`vb_setup_dense`: two Cholesky
        factorisations per block, which also certify that the reference would keep every eigenpair);
        blocks the kernel declines, truncated or factor-form blocks take the exact host path (eigh,
        pseudo-inverse, Woodbury -- the reference's own LAPACK calls) on the set-up thread pool.
        Returns dict(mle, rmle, ridge, chi, rank, gpu_blocks, host_blocks) in this operator's SNP order."""
        from ._pool import map_blocks
        n_tot = self.shape[0]
        z = np.asarray(z_scores, dtype=np.float64)[self.perm]
        reg = np.zeros(n_tot)
        reg[:] = regularizer
        reg = reg[self.perm]
        out = {k: np.zeros(n_tot) for k in ('mle', 'rmle', 'ridge')}
        chi, rank = 0.0, 0
        bounds = list(zip(self.starts[:-1], self.starts[1:]))
        gpu_ids = [b for b, m in enumerate(self.matrices)
                   if m._X is not None and not m.factorized and np.all(m.D == 0)]
        done = set()
        if gpu_ids:
            done = self._device_setup_blocks(gpu_ids, bounds, z, reg, out, ctx)
            for b in done:
                lo, hi = bounds[b]
                chi += float(z[lo:hi].dot(out['mle'][lo:hi]))
                rank += hi - lo
        rest = [b for b in range(len(self.matrices)) if b not in done]

        def host_one(b):
            m = self.matrices[b]
            lo, hi = bounds[b]
            mle = m.inverse_dot(z[lo:hi])
            rmle = m._host_dot(mle)
            shifted = LowRankMatrix(u=m.u, s=m.s, v=m.v, D=m.D + reg[lo:hi])
            return mle, rmle, shifted.inverse_dot(rmle), m.get_rank()
        for b, (mle, rmle, ridge, r) in zip(rest, map_blocks(host_one, rest)):
            lo, hi = bounds[b]
            out['mle'][lo:hi], out['rmle'][lo:hi], out['ridge'][lo:hi] = mle, rmle, ridge
            chi += float(z[lo:hi].dot(mle))
            rank += r
        res = {k: v[self.inv_perm] for k, v in out.items()}
        res.update(chi=chi, rank=rank, gpu_blocks=len(done), host_blocks=len(rest))
        return res

    def _device_setup_blocks(self, ids, bounds, z, reg, out, ctx, chunk_bytes=1 << 30):
        """Run vb_setup_dense over the lazily loaded dense blocks `ids`, in chunks of ~1 GB of matrix
        data; fills `out` for the blocks the kernel certified and returns their ids."""
        import ctypes as C
        import torch
        from . import _lib
        from .engine import DeviceContext
        if ctx is None:
            ctx = DeviceContext.get()
        lib = ctx.lib
        nmax = int(lib.vb_setup_nmax())
        ids = [b for b in ids if self.matrices[b].shape[0] <= nmax]
        dev = torch.device('cuda', ctx.device)
        done = set()
        i = 0
        while i < len(ids):
            chunk, nbytes = [], 0
            while i < len(ids) and (not chunk or nbytes + 8 * self.matrices[ids[i]].shape[0]**2 <= chunk_bytes):
                chunk.append(ids[i])
                nbytes += 8 * self.matrices[ids[i]].shape[0]**2
                i += 1
            ns = np.array([self.matrices[b].shape[0] for b in chunk], dtype=np.int64)
            host = torch.empty(int((ns**2).sum()), dtype=torch.float64, pin_memory=True)
            hv = host.numpy()
            off = 0
            for b, n in zip(chunk, ns):
                hv[off:off + n * n] = self.matrices[b].symmetric_matrix().reshape(-1)
                off += n * n
            R = host.to(dev, non_blocking=True)
            W = torch.empty_like(R)
            zc = torch.as_tensor(np.concatenate([z[bounds[b][0]:bounds[b][1]] for b in chunk]), device=dev)
            rc = torch.as_tensor(np.concatenate([reg[bounds[b][0]:bounds[b][1]] for b in chunk]), device=dev)
            mle, rmle, ridge = (torch.zeros_like(zc) for _ in range(3))
            chi = torch.zeros(len(chunk), dtype=torch.float64, device=dev)
            lam = torch.zeros(2 * len(chunk), dtype=torch.float64, device=dev)
            status = torch.ones(len(chunk), dtype=torch.int32, device=dev)
            ptr = lambda t: C.c_void_p(t.data_ptr())
            _lib.check(lib.vb_setup_dense(ctx.handle, len(chunk), ns.ctypes.data_as(_lib.c_i64p), ptr(R),
                                          ptr(W), ptr(zc), ptr(rc), ptr(mle), ptr(rmle), ptr(ridge),
                                          ptr(chi), ptr(lam), ptr(status)))
            st = status.cpu().numpy()
            mle, rmle, ridge = mle.cpu().numpy(), rmle.cpu().numpy(), ridge.cpu().numpy()
            off = 0
            for k, (b, n) in enumerate(zip(chunk, ns)):
                if st[k] == 0:
                    lo, hi = bounds[b]
                    out['mle'][lo:hi] = mle[off:off + n]
                    out['rmle'][lo:hi] = rmle[off:off + n]
                    out['ridge'][lo:hi] = ridge[off:off + n]
                    self.matrices[b].full_rank_certified = True
                    done.add(b)
                off += n
            del R, W
        return done

    def restrict(self, snps):
        """The operator of the LD blocks that live entirely inside the sorted global SNP set `snps`
        (one rank's shard), renumbered to positions in `snps`.  Blocks must not straddle the set."""
        from .partition import local_blocks
        snps = np.asarray(snps, dtype=np.int64)
        ids, perm_local = local_blocks(self, snps, self.shape[0])
        covered = np.zeros(len(snps), dtype=bool)
        covered[perm_local] = True
        missing = np.where(~covered)[0]
        return BlockDiagonalMatrix([self.matrices[b] for b in ids], inverse=self._inverted,
                                   perm=np.concatenate([perm_local, missing]), missing=missing)

    def release_device(self):
        for d in self._device.values():
            d.close()
        self._device = {}

    # ---- operators ---------------------------------------------------------------
    def dot(self, vector, ctx=None):
        """Matrix @ vector.  Non-inverted: GPU mat-vec.  Inverted: host pseudo-inverse (setup)."""
        vector = np.asarray(vector, dtype=np.float64)
        if self._inverted:
            from ._pool import map_blocks
            xp = vector[self.perm]
            parts = map_blocks(lambda t: t[0].inverse_dot(xp[t[1]:t[2]]),
                               zip(self.matrices, self.starts[:-1], self.starts[1:]))
            parts.append(np.zeros([self.missing.shape[0]] + list(vector.shape[1:])))
            return np.concatenate(parts, axis=0)[self.inv_perm]
        dev = self.to_device(ctx)
        if vector.ndim == 1:
            return dev.dot(vector)
        cols = [dev.dot(np.ascontiguousarray(vector[:, j])) for j in range(vector.shape[1])]
        return np.stack(cols, axis=1)

    def ridge_inverse_dot(self, vector, regularizer):
        """Inverse(Matrix + diag(regularizer)) @ vector   (matrix_structures.py:349-387)."""
        if self._inverted:
            raise NotImplementedError('ridge_inverse_dot with inverted matrices has not been '
                                      'implemented yet.')
        reg = np.zeros_like(vector)
        reg[:] = regularizer
        reg = reg[self.perm]
        xp = vector[self.perm]
        from ._pool import map_blocks

        def one(t):
            m, lo, hi = t
            shifted = LowRankMatrix(u=m.u, s=m.s, v=m.v, D=m.D + reg[lo:hi])
            return shifted.inverse_dot(xp[lo:hi])
        parts = map_blocks(one, zip(self.matrices, self.starts[:-1], self.starts[1:]))
        parts.append(np.zeros(self.missing.shape[0]))
        return np.concatenate(parts, axis=0)[self.inv_perm]

    def matrix_power(self, power):
        return BlockDiagonalMatrix([x.matrix_power(power) for x in self.matrices],
                                   inverse=self._inverted, missing=self.missing)

    @property
    def inverse(self):
        return BlockDiagonalMatrix(self.matrices, inverse=not self._inverted,
                                   perm=self.perm, missing=self.missing)

    def diag(self):
        if self._inverted:
            raise NotImplementedError('Getting the diagonal of an inverted matrix has not '
                                      'been implemented yet.')
        parts = [m.diag() for m in self.matrices]
        parts.append(np.zeros(self.missing.shape[0]))
        return np.concatenate(parts, axis=0)[self.inv_perm]

    def get_rank(self):
        return sum(m.get_rank() for m in self.matrices)
