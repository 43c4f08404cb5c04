"""Per-block set-up work on a host thread pool.

`vilma fit` on 1.2M SNPs spends ~5 s in the fitting loop on a B200 and, in the reference's order
of operations, minutes before it: one `eigh` per LD block at load (load.py:237-354 ->
matrix_structures.py:15-28), one pseudo-inverse product and one Woodbury ridge solve per block in the
constructor (variational_inference.py:236-252 -> matrix_structures.py:159-196, :349-387).  The blocks
are independent and LAPACK releases the GIL, so they run concurrently here, each with single-threaded
BLAS: a 700-SNP block's `eigh` parallelises poorly inside (90 ms with 8 BLAS threads) but eight of
them run side by side at 19 ms per block.  Results are the same LAPACK calls on the same inputs
(differences from the BLAS thread count are at the 1e-15 level, as in the reference itself).

VILMA_B200_SETUP_THREADS=1 restores the serial order of execution.
"""
import os
from concurrent.futures import ThreadPoolExecutor


def setup_threads():
    env = os.environ.get('VILMA_B200_SETUP_THREADS')
    if env:
        return max(1, int(env))
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        cpus = os.cpu_count() or 1
    # ranks of one node load side by side
    local_world = int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1)
    return max(1, min(16, cpus // max(local_world, 1)))


class _SingleThreadBlas:
    def __enter__(self):
        self._ctl = None
        try:
            from threadpoolctl import threadpool_limits
            self._ctl = threadpool_limits(limits=1)
        except Exception:
            pass
        return self

    def __exit__(self, *exc):
        if self._ctl is not None:
            self._ctl.restore_original_limits()
        return False


MIN_BLOCKS = 16     # fewer items than this are not worth a pool (and keep small runs single-threaded)


def map_blocks(fn, items, workers=None):
    """[fn(x) for x in items], evaluated on the pool; order preserved; exceptions propagate."""
    items = list(items)
    workers = setup_threads() if workers is None else workers
    if workers <= 1 or len(items) < max(MIN_BLOCKS, 2):
        return [fn(x) for x in items]
    with _SingleThreadBlas(), ThreadPoolExecutor(workers) as pool:
        return list(pool.map(fn, items))


class OrderedPipeline:
    """submit(fn, *args) as items become available, results() in submission order, with at most
    `depth` items in flight (bounds the memory held by queued inputs)."""

    def __init__(self, workers=None, depth=None):
        self.workers = setup_threads() if workers is None else workers
        self.depth = depth or 2 * self.workers
        self._pool = None
        self._blas = None
        self._pending = []
        self._done = []
        if self.workers > 1:
            self._blas = _SingleThreadBlas().__enter__()
            self._pool = ThreadPoolExecutor(self.workers)

    def submit(self, fn, *args):
        if self._pool is None:
            self._done.append(fn(*args))
            return
        self._pending.append(self._pool.submit(fn, *args))
        while len(self._pending) > self.depth:
            self._done.append(self._pending.pop(0).result())

    def results(self):
        try:
            for fut in self._pending:
                self._done.append(fut.result())
            self._pending = []
            return self._done
        finally:
            self.close()

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        if self._blas is not None:
            self._blas.__exit__(None, None, None)
            self._blas = None
