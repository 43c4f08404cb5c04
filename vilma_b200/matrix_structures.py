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
    """Low rank plus diagonal representation of a symmetric block."""

    def __init__(self, X=None, t=1.0, u=None, s=None, v=None, D=None, hdf_file=None):
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
        return np.einsum('ik,ki->i', self.u * self.s, self.v) + self.D

    def matrix_power(self, power):
        if not np.allclose(self.D, 0):
            raise NotImplementedError('Matrix powers with a non-zero diagonal part are not '
                                      'implemented.')
        return LowRankMatrix(u=self.u, s=self.s**power, v=self.v, D=self.D)

    def get_rank(self):
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
        from .engine import choose_storage, sym_nmax
        sym_nmax()          # load the library on this thread before the workers ask for it
        ids = range(len(self.matrices)) if block_ids is None else block_ids
        return map_blocks(lambda b: self._device_block(self.matrices[b], choose_storage), ids)

    @staticmethod
    def _device_block(m, choose_storage):
        if not np.all(m.D == 0):
            raise NotImplementedError('device LD blocks must have a zero diagonal part D')
        n, r = m.u.shape
        if m.s.shape[0] == 1 and m.s[0] == 0:
            # rank-0 dummy block (matrix_structures.py:141-145): the zero matrix
            return {'n': n, 'kind': 'factor', 'U': np.zeros((n, 1)), 's': np.zeros(1)}
        if choose_storage(n, r) == 'dense':
            return {'n': n, 'kind': 'dense', 'R': (m.u * m.s).dot(m.v)}
        # the reference multiplies by v (= u^T for every block built from X); keep its semantics
        # exactly for a caller-supplied v by folding v into the factor only when it is the
        # transpose, else fall back to the dense product.
        if m.v.shape == m.u.T.shape and np.array_equal(m.v, m.u.T):
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
