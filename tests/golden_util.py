"""Helpers to read the golden fixtures in tests/golden/ (written by tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith('.json'):
        with open(path) as f:
            return json.load(f)
    return np.load(path, allow_pickle=False)


def jstr(z, key):
    return json.loads(str(z[key]))


def csr_arrays(z, prefix):
    return (tuple(int(s) for s in z[prefix + '.shape']), z[prefix + '.indptr'], z[prefix + '.indices'], z[prefix + '.data'])


def coo_arrays(z, prefix):
    return (tuple(int(s) for s in z[prefix + '.shape']), z[prefix + '.row'], z[prefix + '.col'], z[prefix + '.data'])


def monomial_from_coo(z, prefix):
    """(perm, scale) with A[r, perm[r]] = scale[r]; asserts the key has exactly one entry per row."""
    (shape, row, col, data) = coo_arrays(z, prefix)
    assert len(row) == shape[0] and len(np.unique(row)) == shape[0], 'not a monomial key'
    order = np.argsort(row, kind='stable')
    return (col[order].astype(np.int64), data[order].astype(np.float32))


def digest(indptr, indices, data):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(indptr, dtype='<i8').tobytes())
    h.update(np.ascontiguousarray(indices, dtype='<i4').tobytes())
    h.update(np.ascontiguousarray(data, dtype='<f4').tobytes())
    return h.hexdigest()
