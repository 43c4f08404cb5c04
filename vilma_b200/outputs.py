"""Streamed `.npz` output of a fit.

The reference ends `vilma fit` with ``np.savez(<out>, vi_mu, vi_delta, hyper_delta,
error_scaling, scalings, vi_sigma)`` (/root/reference/src/vilma/vi_options.py:263-265), where
``vi_sigma[K,P,P,M]`` (variational_inference.py:712-724) is 8 K P^2 M bytes: 7.5 GB for
BASELINE configs[2], 61 GB for configs[4] -- more than the host holds.  Here the member
``vi_sigma.npy`` is written slice by slice: ``vb_fit_vi_sigma(k0, k1)`` recomputes
S_ki = (Prec_k + diag(sld_i / tau))^-1 on the device for a range of components, a producer thread
lands the slice in one of two page-locked buffers while the main thread writes the other one into
the zip stream, so device compute, the device->host copy and the disk write overlap and the host
never holds more than two slices.  Same member names, shapes, dtypes and byte layout as
``np.savez`` (ZIP_STORED, ZIP64, .npy format 1.0/2.0 headers): ``np.load`` reads it unchanged.
"""
import queue
import threading
import zipfile

import numpy as np
from numpy.lib import format as npy_format

SLICE_BYTES = 256 << 20          # target size of one streamed slice


def _write_header(fid, shape, dtype=np.float64):
    d = {'descr': npy_format.dtype_to_descr(np.dtype(dtype)), 'fortran_order': False,
         'shape': tuple(int(s) for s in shape)}
    try:                                               # format 1.0 unless the header is too long, like np.save
        npy_format.write_array_header_1_0(fid, d)
    except ValueError:
        npy_format.write_array_header_2_0(fid, d)


def _write_array(zipf, name, arr):
    arr = np.asanyarray(arr)
    with zipf.open(name + '.npy', 'w', force_zip64=True) as fid:
        npy_format.write_array(fid, arr, allow_pickle=False)


def slice_plan(K, P, M, slice_bytes=SLICE_BYTES):
    """Component ranges [(k0, k1), ...] of at most `slice_bytes` each (at least one component)."""
    per_k = 8 * P * P * M
    step = max(1, int(slice_bytes // max(per_k, 1)))
    return [(k0, min(K, k0 + step)) for k0 in range(0, K, step)]


def save_fit_npz(path, vi, params, slice_bytes=SLICE_BYTES, stats=None):
    """Write `<path>.npz` (np.savez appends the suffix; so does this) with the members of
    vi_options.py:263-265.  `vi` is the fitted MultiPopVI; only rank 0 writes, every rank takes part
    in gathering the slices.  `stats` (dict, optional) receives {'slices', 'max_slice_bytes'}."""
    if not path.endswith('.npz'):
        path = path + '.npz'
    rank0 = vi._comm.rank == 0
    small = vi.create_dump_dict(params)
    K, P, M = vi.num_mix, vi.num_pops, vi.num_loci
    plan = slice_plan(K, P, M, slice_bytes)
    vi._eng.set_tau(vi.error_scaling)

    # Two page-locked slice buffers cycle between a producer thread (device kernel + device->host
    # copy; ctypes releases the GIL) and this thread (zip stream write; file writes release it too).
    # With several ranks the slices are gathered with collectives, which stay on this thread.
    per_k = P * P * len(vi._snps)
    kmax = max(k1 - k0 for k0, k1 in plan)
    threaded = vi._comm.world == 1
    buffers = [vi._eng._host_array((kmax * per_k,)) for _ in range(2 if threaded else 1)]
    free_q, ready_q = queue.Queue(), queue.Queue()
    for i in range(len(buffers)):
        free_q.put(i)

    def fill(i, k0, k1):
        out = buffers[i][:(k1 - k0) * per_k].reshape(k1 - k0, P, P, len(vi._snps))
        return vi._covariance_slice(k0, k1, out=out)

    def produce():
        try:
            for k0, k1 in plan:
                i = free_q.get()
                ready_q.put((i, fill(i, k0, k1)))
        except BaseException as exc:       # surface device errors in the writing thread
            ready_q.put((None, exc))

    th = None
    if threaded:
        th = threading.Thread(target=produce, name='vi_sigma-producer', daemon=True)
        th.start()
    zipf = zipfile.ZipFile(path, mode='w', compression=zipfile.ZIP_STORED, allowZip64=True) if rank0 else None
    max_bytes = 0
    fid = None
    try:
        if rank0:
            for key, val in small.items():
                _write_array(zipf, key, val)
            fid = zipf.open('vi_sigma.npy', 'w', force_zip64=True)
            _write_header(fid, (K, P, P, M))
        for k0, k1 in plan:
            i, item = ready_q.get() if threaded else (0, fill(0, k0, k1))
            if isinstance(item, BaseException):
                raise item
            max_bytes = max(max_bytes, item.nbytes)
            if rank0:
                fid.write(memoryview(np.ascontiguousarray(item)).cast('B'))
            free_q.put(i)
    finally:
        if fid is not None:
            fid.close()
        if zipf is not None:
            zipf.close()
        if th is not None:
            th.join(timeout=60)
    if stats is not None:
        stats.update(slices=len(plan), max_slice_bytes=int(max_bytes))
    return path
