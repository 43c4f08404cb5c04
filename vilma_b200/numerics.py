"""Small host-side helpers used ONLY at set-up time (constructor / initialisation).

The reference's ``numerics.py`` holds 18 numba kernels; on the fitting loop every one of
them is replaced by the fused CUDA kernels in ``csrc/snp_kernels.cuh`` (see the mapping in
DESIGN.md).  What remains here is the O(K P^2) / O(A K) host arithmetic the constructor
needs before any device state exists (``MultiPopVI.__init__`` :622-626) and the one-off,
seed-dependent ``_initialize`` (:643-700) that must consume NumPy's legacy RNG stream exactly
as the reference does.
"""
import numpy as np

EPSILON = 1e-100    # numerics.py:8


def small_inverse(mats):
    """Inverse of a [K,P,P] stack (mixture covariances -> precisions, numerics.py:238-254)."""
    P = mats.shape[-1]
    if P == 1:
        return 1. / mats
    if P == 2:
        a, b, c, d = mats[:, 0, 0], mats[:, 0, 1], mats[:, 1, 0], mats[:, 1, 1]
        idet = 1. / (a * d - b * c)
        out = np.empty_like(mats)
        out[:, 0, 0], out[:, 1, 1] = d * idet, a * idet
        out[:, 1, 0] = -c * idet
        out[:, 0, 1] = out[:, 1, 0]
        return out
    return np.linalg.inv(mats)


def small_log_det(mats):
    """log-determinant of a [K,P,P] stack (numerics.py:274-290)."""
    P = mats.shape[-1]
    if P == 1:
        return np.log(mats[:, 0, 0])
    if P == 2:
        return np.log(mats[:, 0, 0] * mats[:, 1, 1] - mats[:, 0, 1] * mats[:, 1, 0])
    return np.linalg.slogdet(mats)[1]


def vi_delta_grad_table(hyper_delta, log_det):
    """[A,K-1] table t[a,k] = (log h_ak - ld_k/2) - (log h_aK - ld_K/2); the reference's
    fast_vi_delta_grad (numerics.py:149-164) is this table indexed by annotation."""
    t = np.log(hyper_delta) - 0.5 * log_det
    return t[:, :-1] - t[:, -1:]


def sum_annotations_host(deltas, annotations, num_annotations):
    """numerics.py:118-129 on the host (initialisation only)."""
    out = np.zeros((num_annotations, deltas.shape[1]))
    for a in range(num_annotations):
        out[a] = deltas[annotations == a].sum(axis=0)
    return out
