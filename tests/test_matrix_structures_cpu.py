"""Host half of the LD operator classes (vilma_b200.matrix_structures) against dense NumPy algebra.

Mirrors what the reference checks for its own classes (tests/test.py:28-477: _svd_threshold,
LowRankMatrix init / inverse_dot / diag / matrix_power / get_rank, BlockDiagonalMatrix init /
ridge_inverse_dot / inverse / diag / get_rank / matrix_power) for the methods that stay on the host --
the set-up side of the fit (variational_inference.py:226-252).  The mat-vec itself (`dot`) is the GPU
operator: tests/test_gpu_parity.py::test_ld_dot.
"""
import numpy as np
import pytest

from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix, _svd_threshold


def spd(n, rng, shift=3.0):
    x = rng.random((n, n))
    return x + x.T + shift * np.eye(n)


def dense(m):
    return (m.u * m.s) @ m.v + np.diag(m.D)


def test_svd_threshold_keeps_eigenvalues_above_cut():
    rng = np.random.default_rng(0)
    x = spd(6, rng)
    u, s, v = _svd_threshold(x, 1.0)
    assert np.allclose((u * s) @ v, x)
    for t in np.linspace(0, 1, 21):
        u, s, v = _svd_threshold(x, t)
        assert np.all(s >= 1 - np.sqrt(t)) or (s.shape == (1,) and s[0] == 0)
    x = np.eye(5)
    x[0, 0] = 0
    u, s, v = _svd_threshold(x, 0.5)
    assert s.shape[0] == 4 and np.allclose((u * s) @ v, x)
    # nothing survives the cut: the rank-0 stand-in
    u, s, v = _svd_threshold(0.01 * np.eye(3), 0.5)
    assert u.shape == (3, 1) and s.shape == (1,) and s[0] == 0 and v.shape == (1, 3)


def test_low_rank_matrix_construction_and_errors():
    rng = np.random.default_rng(1)
    with pytest.raises(ValueError):
        bad = np.eye(5)
        bad[0, 1] = 2
        LowRankMatrix(X=bad)
    with pytest.raises(ValueError):
        LowRankMatrix(X=np.eye(5), u=3)
    with pytest.raises(ValueError):
        LowRankMatrix()
    x = spd(5, rng)
    m = LowRankMatrix(X=x, t=1.0)
    assert m.shape == (5, 5) and np.allclose(dense(m), x)
    assert np.allclose(m.inv_s, 1.0 / m.s) and np.all(m.D == 0)
    u, s, v = np.linalg.svd(x)
    m2 = LowRankMatrix(u=u, s=s, v=v, D=np.zeros(5))
    assert np.allclose(dense(m2), x) and m2.shape == (5, 5)
    # a singular block keeps only its non-zero directions
    x = np.eye(5)
    x[0, 0] = 0
    m3 = LowRankMatrix(X=x)
    assert m3.s.shape[0] == 4 and m3.u.shape == (5, 4) and m3.v.shape == (4, 5)
    # factors with tiny singular values are cut at 1e-12 of the largest (reference :119)
    m4 = LowRankMatrix(u=np.eye(3), s=np.array([1.0, 1e-14, 0.5]), v=np.eye(3), D=np.zeros(3), t=1.0)
    assert m4.s.tolist() == [1.0, 0.5]


def test_low_rank_matrix_inverse_dot_three_branches():
    rng = np.random.default_rng(2)
    x = spd(7, rng)
    v = rng.normal(size=7)
    # D == 0: pseudo-inverse through the factors (full rank here: the inverse)
    m = LowRankMatrix(X=x)
    assert np.allclose(m.inverse_dot(v), np.linalg.solve(x, v))
    # rank-deficient, D == 0: the Moore-Penrose solution
    g = rng.normal(size=(3, 7))
    low = g.T @ g
    ml = LowRankMatrix(X=low / np.abs(low).max(), t=1.0)
    assert np.allclose(ml.inverse_dot(v), np.linalg.pinv(dense(ml), rcond=1e-10) @ v, atol=1e-8)
    # D > 0 everywhere: Woodbury
    d = rng.uniform(0.5, 2.0, size=7)
    mw = LowRankMatrix(u=m.u, s=m.s, v=m.v, D=d)
    assert np.allclose(mw.inverse_dot(v), np.linalg.solve(x + np.diag(d), v))
    # mixed zero / non-zero D: dense pseudo-inverse with the data-dependent rcond
    d2 = d.copy()
    d2[:3] = 0
    mm = LowRankMatrix(u=m.u, s=m.s, v=m.v, D=d2)
    assert np.allclose(mm.inverse_dot(v), np.linalg.solve(x + np.diag(d2), v), rtol=1e-8)


def test_low_rank_matrix_diag_power_rank():
    rng = np.random.default_rng(3)
    x = spd(6, rng)
    m = LowRankMatrix(X=x)
    assert np.allclose(m.diag(), np.diag(x))
    d = rng.uniform(0.1, 1.0, size=6)
    md = LowRankMatrix(u=m.u, s=m.s, v=m.v, D=d)
    assert np.allclose(md.diag(), np.diag(x) + d)
    half = m.matrix_power(0.5)
    assert np.allclose(dense(half) @ dense(half), x)
    inv = m.matrix_power(-1)
    assert np.allclose(dense(inv), np.linalg.inv(x))
    with pytest.raises(NotImplementedError):
        md.matrix_power(2)
    assert m.get_rank() == 6 and md.get_rank() == 6
    e = np.eye(5)
    e[0, 0] = 0
    assert LowRankMatrix(X=e).get_rank() == 4
    assert LowRankMatrix(X=0.01 * np.eye(3), t=0.5).get_rank() == 0       # rank-0 stand-in
    one = LowRankMatrix(u=np.ones((3, 1)) / np.sqrt(3), s=np.array([2.0]), v=np.ones((1, 3)) / np.sqrt(3),
                        D=np.zeros(3))
    assert one.get_rank() == 1
    mixed = LowRankMatrix(u=one.u, s=one.s, v=one.v, D=np.array([0.0, 0.0, 1.0]))
    assert mixed.get_rank() == np.linalg.matrix_rank(dense(mixed))


def make_bdm(rng, sizes=(4, 6, 3), n_missing=2, shuffle=True):
    mats = [spd(n, rng) for n in sizes]
    n = sum(sizes) + n_missing
    perm = rng.permutation(n) if shuffle else np.arange(n)
    missing = perm[sum(sizes):]
    bdm = BlockDiagonalMatrix([LowRankMatrix(X=m) for m in mats], perm=perm, missing=missing)
    full = np.zeros((n, n))
    lo = 0
    for m in mats:
        idx = perm[lo:lo + m.shape[0]]
        full[np.ix_(idx, idx)] = m
        lo += m.shape[0]
    return bdm, full, missing


def test_block_diagonal_construction_and_errors():
    rng = np.random.default_rng(4)
    with pytest.raises(ValueError):
        BlockDiagonalMatrix([np.eye(3)])
    blocks = [LowRankMatrix(X=spd(3, rng)), LowRankMatrix(X=spd(2, rng))]
    with pytest.raises(ValueError):
        BlockDiagonalMatrix(blocks, perm=np.arange(4))                       # wrong length
    with pytest.raises(ValueError):
        BlockDiagonalMatrix(blocks, perm=np.array([0, 1, 2, 3, 3]))          # not a permutation
    bdm = BlockDiagonalMatrix(blocks, missing=np.array([5, 6]))
    assert bdm.shape == (7, 7) and np.array_equal(bdm.perm, np.arange(7))
    assert np.array_equal(bdm.starts, [0, 3, 5])
    bdm2, _, missing = make_bdm(rng)
    assert np.array_equal(bdm2.perm[bdm2.inv_perm], np.arange(bdm2.shape[0]))
    assert np.array_equal(np.sort(bdm2.perm[int(bdm2.starts[-1]):]), np.sort(missing))


def test_block_diagonal_inverse_ridge_diag_rank_power():
    rng = np.random.default_rng(5)
    bdm, full, missing = make_bdm(rng)
    n = full.shape[0]
    have = np.setdiff1d(np.arange(n), missing)
    z = rng.normal(size=n)
    # pseudo-inverse product: zero on SNPs without LD, block inverses elsewhere
    got = bdm.inverse.dot(z)
    want = np.zeros(n)
    want[have] = np.linalg.solve(full[np.ix_(have, have)], z[have])
    assert np.allclose(got, want) and np.all(got[missing] == 0)
    assert isinstance(bdm.inverse.inverse, BlockDiagonalMatrix) and not bdm.inverse.inverse._inverted
    # ridge solve with a per-SNP regulariser
    reg = rng.uniform(0.1, 1.0, size=n)
    got = bdm.ridge_inverse_dot(z, reg)
    want = np.zeros(n)
    want[have] = np.linalg.solve(full[np.ix_(have, have)] + np.diag(reg[have]), z[have])
    assert np.allclose(got, want) and np.all(got[missing] == 0)
    with pytest.raises(NotImplementedError):
        bdm.inverse.ridge_inverse_dot(z, reg)
    # scalar regulariser broadcasts like the reference's `reg[:] = regularizer`
    assert np.allclose(bdm.ridge_inverse_dot(z, 0.3)[have],
                       np.linalg.solve(full[np.ix_(have, have)] + 0.3 * np.eye(len(have)), z[have]))
    assert np.allclose(bdm.diag(), np.diag(full))
    with pytest.raises(NotImplementedError):
        bdm.inverse.diag()
    assert bdm.get_rank() == len(have)
    half = bdm.matrix_power(0.5)
    assert [m.shape for m in half.matrices] == [m.shape for m in bdm.matrices]
    for a, b in zip(half.matrices, bdm.matrices):
        assert np.allclose(dense(a) @ dense(a), dense(b))
