"""The reference-side binding shown in INTEGRATION.md section 2 is EXECUTED verbatim (only the library path is filled in and a
stand-in for the reference's `keynet.sparse.SparseMatrix` base class is provided -- /root/reference does not exist on the
GPU box): a scipy CSR matrix goes in, `torchdot` runs on the B200 through the C ABI, the result is compared with the oracle."""
import os
import re
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_runs_against_scipy_csr():
    scipy_sparse = pytest.importorskip('scipy.sparse')
    from keynet_b200 import _native
    from oracle import keynet_oracle as ko
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = re.search(r"## 2\..*?```python\n(.*?)```", text, re.S).group(1)
    assert '/path/to/keynet_b200/lib/libkeynet_b200.so' in code
    code = code.replace('/path/to/keynet_b200/lib/libkeynet_b200.so', _native.LIB_PATH)
    # stand-in for the reference's operator base class (keynet/sparse.py:419-462): holds the scipy matrix and its shape
    keynet = types.ModuleType('keynet')
    ksparse = types.ModuleType('keynet.sparse')

    class SparseMatrix(object):
        def __init__(self, A=None):
            self._matrix = A
            self.shape = (0, 0) if A is None else A.shape
            self.dtype = None if A is None else A.dtype
            self.ndim = 2
    ksparse.SparseMatrix = SparseMatrix
    keynet.sparse = ksparse
    saved = {k: sys.modules.get(k) for k in ('keynet', 'keynet.sparse')}
    sys.modules['keynet'] = keynet; sys.modules['keynet.sparse'] = ksparse
    try:
        ns = {}
        exec(compile(code, 'INTEGRATION.md#2', 'exec'), ns)
    finally:
        for (k, v) in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rs = np.random.RandomState(0)
    A = scipy_sparse.random(300, 200, density=0.05, random_state=rs, dtype=np.float32, format='coo')
    W = ns['SparseMatrix'](A)
    X = rs.randn(200, 37).astype(np.float32)
    for relu in (False, True):
        y = W.torchdot(torch.from_numpy(X), relu=relu)
        assert not y.is_cuda and tuple(y.shape) == (300, 37)
        C = A.tocsr(); C.sort_indices()
        ref = ko.spmm(ko.csr(C.shape, C.indptr, C.indices, C.data), X, relu=relu)
        assert np.allclose(y.numpy(), ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())
    y = W.torchdot(torch.from_numpy(X).cuda())
    assert y.is_cuda
