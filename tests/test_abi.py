"""The C-ABI library loads on a CPU-only box and exports exactly the symbols include/keynet_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    from keynet_b200 import _native
    return _native


def _header_symbols():
    with open(os.path.join(ROOT, 'include', 'keynet_b200.h')) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(kn_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_agree(built):
    assert _header_symbols() == sorted(built.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    L = built.lib()
    for s in _header_symbols():
        assert hasattr(L, s), s
    assert L.kn_abi_version() == 2


def test_header_cites_reference_interfaces():
    with open(os.path.join(ROOT, 'include', 'keynet_b200.h')) as f:
        src = f.read()
    for cite in ['keynet/sparse.py:488-492', 'keynet/layer.py:35', 'keynet/sparse.py:122-212', 'keynet/torch.py:65-68']:
        assert cite in src


def test_argument_validation_without_gpu(built):
    """Invalid arguments are rejected with an error code and a message before any CUDA work."""
    L = built.lib()
    rc = L.kn_spmm_csr_f32(None, None, None, 4, 4, None, 1, None, 1, 2, 0, None, None)     # ld < n_vecs
    assert rc == -1 and b'leading dimension' in L.kn_last_error()
    rc = L.kn_exclusive_scan_i64(None, None, -1, None)
    assert rc == -1
    bad = built.kn_peers(); bad.n = 9                                                  # more than 8 destinations
    rc = L.kn_spmm_csr_f32(None, None, None, 4, 4, None, 2, None, 2, 2, 0, ctypes.byref(bad), None)
    assert rc == -1 and b'destinations' in L.kn_last_error()
    d = built.kn_conv2d_desc(1, 8, 8, 1, 2, 2, 1, 0, 1)                               # even kernel
    rc = L.kn_toeplitz_conv2d_count(d, None, 1, None, None)
    assert rc == -1 and b'odd' in L.kn_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'keynet_b200')
    for (dirpath, _, files) in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                with open(os.path.join(dirpath, fn)) as f:
                    s = f.read()
                assert 'import oracle' not in s and 'from oracle' not in s and 'keynet_oracle' not in s, fn


def test_no_cpu_fallback_without_cuda(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from keynet_b200 import sparse
    with pytest.raises(built.NativeError):
        sparse.SparseMatrix(sparse.sparse_identity_matrix(4))
