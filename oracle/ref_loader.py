"""Import the UNMODIFIED reference package installed in `oracle/_ref/` (see build_ref.py).

Test / benchmark infrastructure only: `bench.py --impl reference` and the `cpu_baseline` leg time
the reference's own numba + NumPy fit through this loader (`kind: "reference"`); nothing in
`vilma_b200/` imports it.  The same four import-time shims as tests/golden/_ref_shim.py make the
2022-era code run on numba 0.65 / pandas 3 without editing any of its files:

1. ``numerics.sum_annotations`` (numerics.py:118-129) does an array ``+=`` inside ``prange`` and
   corrupts the heap under numba 0.65 -> NumPy ``np.add.at``;
2. ``h5py`` is not installed (only used for ``--mmap``) -> stub module;
3. ``pd.read_csv(delim_whitespace=True)`` was removed in pandas 3 -> ``sep=r'\\s+'``;
4. pandas-3 copy-on-write makes ``Series.to_numpy()`` read-only but load.py:286 writes into it.
"""
import os
import sys
import types

import numpy as np

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    """(ok, reason): can the installed reference be imported and JIT-compiled here?"""
    if not os.path.exists(os.path.join(REF_DIR, 'vilma', 'variational_inference.py')):
        return False, 'oracle/_ref is not installed (python oracle/build_ref.py needs /root/reference)'
    try:
        import numba  # noqa: F401
    except Exception as exc:
        return False, 'numba cannot be imported on this host: %r' % (exc,)
    return True, ''


def import_reference():
    """Namespace(numerics, matrix_structures, variational_inference, load, vi_options) of the
    reference, shimmed.  Raises ImportError when it is not installed or numba is missing."""
    ok, why = available()
    if not ok:
        raise ImportError(why)
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules['h5py'] = types.ModuleType('h5py')
    import pandas as pd
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
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import vilma  # noqa: F401
    from vilma import numerics

    def sum_annotations(deltas, annotations, num_annotations):
        out = np.zeros((num_annotations, deltas.shape[1]))
        np.add.at(out, annotations, deltas)
        return out
    numerics.sum_annotations = sum_annotations
    from vilma import load, matrix_structures, variational_inference, vi_options
    return types.SimpleNamespace(numerics=numerics, matrix_structures=matrix_structures,
                                 variational_inference=variational_inference, load=load,
                                 vi_options=vi_options, path=REF_DIR)
