"""NumPy restatement of the reference LD operator (ORACLE -- tests only).

Follows ``/root/reference/src/vilma/matrix_structures.py``.  A block is held as
``U diag(s) U^T + diag(D)`` (matrix_structures.py:72-146); the block-diagonal
operator gathers ``x[perm]``, applies each block, appends zeros for ``missing``
SNPs and scatters with ``inv_perm`` (matrix_structures.py:389-408).
"""
import numpy as np


def eig_threshold(matrix, ld_thresh):
    """matrix_structures.py:15-28  keep eigenpairs with lambda >= 1 - sqrt(t)"""
    vals, vecs = np.linalg.eigh(matrix)
    keep = np.where(vals >= 1 - np.sqrt(ld_thresh))[0]
    if len(keep) == 0:
        n = matrix.shape[0]
        return np.ones((n, 1)), np.zeros(1), np.ones((1, n))
    u = np.copy(vecs[:, keep])
    return u, np.copy(vals[keep]), np.copy(u.T)


class LowRankBlock:
    """One LD block, U diag(s) V + diag(D)   (matrix_structures.py:38-234)."""

    def __init__(self, X=None, t=1.0, u=None, s=None, v=None, D=None):
        if X is not None:
            if not np.allclose(X, X.T):
                raise ValueError('Provided matrix is not symmetric')
            u, s, v = eig_threshold(X, t)
            D = np.zeros(X.shape[0])
        else:
            sel = np.where(s >= 1 - np.sqrt(t))[0]      # :113-116
            u, s, v = u[:, sel], s[sel], v[sel, :]
        self.D = np.array(D, dtype=float)
        big = s > 1e-12 * np.max(s)                      # :119
        if big.sum() > 0:
            self.u = np.array(u[:, big], dtype=float)
            self.s = np.array(s[big], dtype=float)
            self.v = np.array(v[big, :], dtype=float)
            self.inv_s = 1. / self.s
        else:                                            # :141-145 rank-0 dummy
            self.u = np.array(u[:, :1], dtype=float)
            self.s = np.zeros(1)
            self.v = np.array(v[:1, :], dtype=float)
            self.inv_s = np.zeros(1)
        self.shape = (self.u.shape[0], self.v.shape[1])

    def dot(self, x):
        """:148-152"""
        mid = (self.s * self.v.dot(x).T).T
        return self.u.dot(mid) + (self.D * x.T).T

    def inverse_dot(self, x):
        """:159-196  pseudo-inverse (D==0), pinv (mixed D) or Woodbury (D != 0)"""
        if np.any(np.isclose(np.abs(self.D), 0)):
            if np.all(np.isclose(self.D, 0)):
                return self.v.T.dot(self.u.T.dot(x) * self.inv_s)
            full = np.diag(self.D) + (self.u * self.s).dot(self.v)
            ev = np.linalg.eigh(full)[0][::-1]
            hit = np.where(np.isclose(np.cumsum(ev) / np.sum(ev), 1.))[0]
            pos = hit[0] if len(hit) > 0 else len(ev) - 1
            rcond = ev[pos] / ev[0] * 0.1
            return np.linalg.pinv(full, rcond=rcond).dot(x)
        small = np.diag(self.inv_s) + self.v.dot((self.u.T / self.D).T)
        small = np.linalg.inv(small)
        t = self.u.dot(small.dot(self.v.dot(x / self.D)))
        t /= self.D
        return x / self.D - t

    def diag(self):
        """:198-203"""
        return np.einsum('ik,ki->i', self.u * self.s, self.v) + self.D

    def get_rank(self):
        """:213-234"""
        if np.allclose(self.D, 0):
            if self.s.shape[0] > 1:
                return self.s.shape[0]
            return 0 if self.s[0] == 0 else 1
        if np.all(self.D > 0):
            return self.D.shape[0]
        full = np.diag(self.D) + np.einsum('ik,k,kj->ij', self.u, self.s, self.v)
        return np.linalg.matrix_rank(full, hermitian=True)


class BlockDiagonalLD:
    """Block-diagonal LD with permutation and missing SNPs (:237-447)."""

    def __init__(self, blocks, perm=None, missing=None, inverted=False):
        self.missing = (np.array([], dtype=np.int64) if missing is None
                        else np.array(missing, dtype=np.int64))
        self.matrices = list(blocks)
        self._inverted = inverted
        sizes = [b.shape[0] for b in self.matrices]
        self.starts = np.cumsum([0] + sizes)
        n = int(sum(sizes)) + self.missing.shape[0]
        self.shape = (n, n)
        self.perm = np.arange(n) if perm is None else np.array(perm, dtype=np.int64)
        if self.perm.shape[0] != n:
            raise ValueError('perm must be conformal to the matrix')
        self.inv_perm = np.argsort(self.perm)

    def _blockwise(self, x, fn):
        xp = x[self.perm]
        parts = [fn(b, xp[lo:hi]) for b, lo, hi in
                 zip(self.matrices, self.starts[:-1], self.starts[1:])]
        parts.append(np.zeros([self.missing.shape[0]] + list(x.shape[1:])))
        return np.concatenate(parts, axis=0)[self.inv_perm]

    def dot(self, x):
        """:389-408"""
        if self._inverted:
            return self._blockwise(x, lambda b, v: b.inverse_dot(v))
        return self._blockwise(x, lambda b, v: b.dot(v))

    @property
    def inverse(self):
        """:418-424"""
        return BlockDiagonalLD(self.matrices, perm=self.perm, missing=self.missing,
                               inverted=not self._inverted)

    def ridge_inverse_dot(self, x, regularizer):
        """:349-387  (R + diag(reg))^-1 x, blockwise"""
        reg = np.zeros_like(x)
        reg[:] = regularizer
        reg = reg[self.perm]
        xp = x[self.perm]
        parts = []
        for b, lo, hi in zip(self.matrices, self.starts[:-1], self.starts[1:]):
            shifted = LowRankBlock(u=b.u, s=b.s, v=b.v, D=b.D + reg[lo:hi])
            parts.append(shifted.inverse_dot(xp[lo:hi]))
        parts.append(np.zeros(self.missing.shape[0]))
        return np.concatenate(parts, axis=0)[self.inv_perm]

    def diag(self):
        """:426-440"""
        parts = [b.diag() for b in self.matrices]
        parts.append(np.zeros(self.missing.shape[0]))
        return np.concatenate(parts, axis=0)[self.inv_perm]

    def get_rank(self):
        """:442-447"""
        return sum(b.get_rank() for b in self.matrices)
