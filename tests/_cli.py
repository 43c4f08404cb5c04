"""Write the input files packed in a CLI golden fixture back to disk."""
import io
import os

import numpy as np
import pandas as pd


def materialize(fx, root):
    for key, val in fx.items():
        if not key.startswith('in_'):
            continue
        rel = key[3:]
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path) or root, exist_ok=True)
        if rel.endswith('.npy'):
            np.save(path, val)
        else:
            with open(path, 'w') as fh:
                fh.write(str(val))
    return root


def read_tsv(text):
    return pd.read_csv(io.StringIO(str(text)), sep='\t', header=0)


def frames_close(a, b, rtol=1e-5, atol=1e-8):
    """Like the reference's check_data_frame: same columns, numeric columns allclose."""
    assert list(a.columns) == list(b.columns), (list(a.columns), list(b.columns))
    assert len(a) == len(b)
    for col in a.columns:
        if pd.api.types.is_float_dtype(a[col]) or pd.api.types.is_float_dtype(b[col]):
            assert np.allclose(a[col].to_numpy(dtype=float), b[col].to_numpy(dtype=float),
                               rtol=rtol, atol=atol), col
        else:
            assert (a[col].astype(str) == b[col].astype(str)).all(), col
    return True
