"""Import the UNMODIFIED reference (jeffspence/vilma) from /root/reference/src.

Container-only helper used by ``make_golden.py`` to produce the committed
fixtures in this directory.  /root/reference does not exist on the GPU box, so
nothing in ``tests/``, ``bench.py`` or the package imports this module at run
time.  No reference file is edited or copied; four import-time shims make the
2022-era code run on numba 0.65 / pandas 3:

1. ``numerics.sum_annotations`` (numerics.py:118-129) does an array ``+=`` inside
   ``prange`` and corrupts the heap under numba 0.65 -> NumPy ``np.add.at``.
2. ``h5py`` is not installed (only used for ``--mmap``) -> stub module.
3. ``pd.read_csv(delim_whitespace=True)`` was removed in pandas 3 -> ``sep=r'\\s+'``.
4. pandas-3 copy-on-write makes ``Series.to_numpy()`` read-only but
   load.py:286 writes into it -> hand back a writable copy.
"""
import sys
import types

import numpy as np
import pandas as pd

REF_SRC = '/root/reference/src'


def import_reference():
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules['h5py'] = types.ModuleType('h5py')

    if not getattr(pd.read_csv, '_vilma_shim', False):
        _orig_read_csv = pd.read_csv

        def read_csv(*args, **kwargs):
            if kwargs.pop('delim_whitespace', False):
                kwargs['sep'] = r'\s+'
            return _orig_read_csv(*args, **kwargs)
        read_csv._vilma_shim = True
        pd.read_csv = read_csv

        _orig_to_numpy = pd.Series.to_numpy

        def to_numpy(self, *args, **kwargs):
            out = _orig_to_numpy(self, *args, **kwargs)
            if isinstance(out, np.ndarray) and not out.flags.writeable:
                out = out.copy()
            return out
        pd.Series.to_numpy = to_numpy

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import vilma  # noqa: F401
    from vilma import numerics

    def sum_annotations(deltas, annotations, num_annotations):
        out = np.zeros((num_annotations, deltas.shape[1]))
        np.add.at(out, annotations, deltas)
        return out
    numerics.sum_annotations = sum_annotations

    from vilma import matrix_structures, variational_inference, load, vi_options
    return types.SimpleNamespace(
        numerics=numerics, matrix_structures=matrix_structures,
        variational_inference=variational_inference, load=load,
        vi_options=vi_options)
