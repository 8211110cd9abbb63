"""Fused key compile (csrc/keyedconv.cu): the one-pass CSR writer and the direct pattern-group builder against the two-kernel
path (Toeplitz CSR -> per-row key compile -> pattern hashing), which is itself pinned bit-for-bit to the reference goldens
(tests/test_gpu_parity.py), and against the oracle's csr_matmat compile.  Bars: index arrays and values bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _keys(rs, n_out, n_in, perm_out, perm_in, gain):
    from keynet_b200.sparse import MonomialKey
    po = np.concatenate([rs.permutation(n_out - 1) if perm_out else np.arange(n_out - 1), [n_out - 1]])
    pi = np.concatenate([rs.permutation(n_in - 1) if perm_in else np.arange(n_in - 1), [n_in - 1]])
    a = np.concatenate([rs.rand(n_out - 1) + 0.5, [1.0]]).astype(np.float32) if gain else None
    ai = np.concatenate([rs.rand(n_in - 1) + 0.5, [1.0]]).astype(np.float32) if gain else None
    return (MonomialKey(po, a), MonomialKey(pi, ai))


def _groups_to_csr(W):
    """Expand the pattern-group format back to canonical CSR arrays on the host (exact zeros dropped)."""
    pg = W._pg
    (r_all, c_all, v_all) = ([], [], [])
    for c in pg.classes:
        (G, K_pad, ng) = (c['G'], c['K_pad'], c['n_groups'])
        rows = c['rows'].cpu().numpy().reshape(ng, G)
        cols = c['cols'].cpu().numpy().reshape(ng, K_pad)
        gk = c['group_k'].cpu().numpy()
        vals = c['vals'].cpu().numpy().reshape(-1, G, K_pad)
        blk = c['block_of'].cpu().numpy() if c['block_of'] is not None else np.arange(ng)
        for g in range(ng):
            k = int(gk[g])
            v = vals[blk[g]][:, :k]
            r_all.append(np.repeat(rows[g], k)); c_all.append(np.tile(cols[g, :k], G)); v_all.append(v.reshape(-1))
    if pg.rest is not None:
        ip = pg.rest['indptr'].cpu().numpy()
        r_all.append(np.repeat(pg.rest['out_rows'].cpu().numpy(), np.diff(ip))); c_all.append(pg.rest['indices'].cpu().numpy()); v_all.append(pg.rest['data'].cpu().numpy())
    (r, c, v) = (np.concatenate(r_all).astype(np.int64), np.concatenate(c_all).astype(np.int64), np.concatenate(v_all).astype(np.float32))
    keep = v != 0
    (r, c, v) = (r[keep], c[keep], v[keep])
    o = np.lexsort((c, r))
    indptr = np.zeros(W.shape[0] + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    return (np.cumsum(indptr), c[o].astype(np.int32), v[o])


def _same(a, b, what):
    assert np.array_equal(a[0] - a[0][0], b[0] - b[0][0]), what + ' indptr'
    assert np.array_equal(a[1], b[1]), what + ' indices'
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)), what + ' data bits'


CASES = [((3, 8, 8), 16, 3, 1), ((16, 12, 12), 8, 3, 2), ((1, 28, 28), 6, 3, 1), ((6, 14, 14), 16, 5, 1), ((8, 6, 6), 12, 1, 1), ((96, 8, 8), 96, 3, 1)]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('keys', [(False, True, False), (True, True, False), (True, True, True), (False, False, False)])
def test_fused_compile_equals_two_kernel_path(case, keys):
    from keynet_b200 import sparse
    ((C, U, V), M, k, stride) = case
    rs = np.random.RandomState(11)
    f = rs.randn(M, C, k, k).astype(np.float32)
    f[rs.rand(*f.shape) < 0.05] = 0
    b = rs.randn(M).astype(np.float32)
    (R, K) = (M * (U // stride) * (V // stride) + 1, C * U * V + 1)
    (A, Ainv) = _keys(rs, R, K, *keys)
    try:
        sparse.direct_compile_enabled(False)
        ref = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, stride, A, Ainv)
    finally:
        sparse.direct_compile_enabled(True)
    got = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, stride, A, Ainv)
    assert got.shape == ref.shape
    _same(got.csr_arrays(), ref.csr_arrays(), 'fused CSR')
    if got._pg is not None:
        _same(_groups_to_csr(got), ref.csr_arrays(), 'direct groups')
    # groups only (the VGG16 route): same groups, same nnz, no CSR
    only = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, stride, A, Ainv, want_csr=False)
    if only._data is None:
        assert only.nnz() == ref.nnz()
        _same(_groups_to_csr(only), ref.csr_arrays(), 'groups without CSR')
    # the product through the groups equals the product through the reference CSR
    X = torch.randn(K, 64, device='cuda')
    try:
        sparse.direct_compile_enabled(False)
        y_ref = sparse.spmm(sparse.SparseMatrix((ref.shape, *ref.csr_arrays())), X, relu=True)       # CSR kernel, no groups
    finally:
        sparse.direct_compile_enabled(True)
    y = sparse.spmm(only, X, relu=True)
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-5 * float(y_ref.abs().max()))


def test_fused_compile_row_shard_and_remap():
    """A shard that owns whole pixels (dist.plan_rows) with a gathered column layout."""
    from keynet_b200 import sparse
    rs = np.random.RandomState(5)
    ((C, U, V), M) = ((6, 14, 14), 16)
    f = rs.randn(M, C, 3, 3).astype(np.float32); b = rs.randn(M).astype(np.float32)
    (R, K) = (M * U * V + 1, C * U * V + 1)
    (A, Ainv) = _keys(rs, R, K, True, True, True)
    inv = np.empty(R, dtype=np.int64); inv[A.perm] = np.arange(R)
    pix = np.arange(40, 90)
    rows = inv[(np.arange(M).reshape(1, -1) * U * V + pix.reshape(-1, 1)).reshape(-1)]         # pixel-major, like plan_rows
    n_phys = K + 11
    remap = np.concatenate([rs.permutation(n_phys - 1)[:K - 1], [n_phys - 1]]).astype(np.int64)
    try:
        sparse.direct_compile_enabled(False)
        ref = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, 1, A, Ainv, rows=rows, col_remap=remap, n_cols_phys=n_phys)
    finally:
        sparse.direct_compile_enabled(True)
    got = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, 1, A, Ainv, rows=rows, col_remap=remap, n_cols_phys=n_phys)
    _same(got.csr_arrays(), ref.csr_arrays(), 'sharded fused CSR')
    assert got._pg is not None and got._pg.rest is None
    _same(_groups_to_csr(got), ref.csr_arrays(), 'sharded direct groups')


@pytest.mark.parametrize('gain', [False, True])
def test_fused_linear_equals_two_kernel_path_and_oracle(gain):
    from keynet_b200 import sparse
    from oracle import keynet_oracle as ko
    rs = np.random.RandomState(2)
    (n_out, n_in) = (120, 784)
    w = rs.randn(n_out, n_in).astype(np.float32); w[rs.rand(*w.shape) < 0.02] = 0
    b = rs.randn(n_out).astype(np.float32)
    (A, Ainv) = _keys(rs, n_out + 1, n_in + 1, True, True, gain)
    got = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv)
    try:
        sparse.direct_compile_enabled(False)
        ref = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv)
    finally:
        sparse.direct_compile_enabled(True)
    _same(got.csr_arrays(), ref.csr_arrays(), 'fused linear CSR')
    _same(_groups_to_csr(got), ref.csr_arrays(), 'direct linear groups')
    o = ko.sort_indices(ko.key_compile(ko.monomial_key(A.perm, A.scale), ko.linear_matrix(w, b), ko.monomial_key(Ainv.perm, Ainv.scale)))
    _same(got.csr_arrays(), (o.indptr, o.indices, o.data), 'oracle')


def test_fused_conv_equals_oracle_compile():
    from keynet_b200 import sparse
    from oracle import keynet_oracle as ko
    rs = np.random.RandomState(4)
    ((C, U, V), M) = ((3, 10, 10), 8)
    f = rs.randn(M, C, 3, 3).astype(np.float32); b = rs.randn(M).astype(np.float32)
    (R, K) = (M * U * V + 1, C * U * V + 1)
    (A, Ainv) = _keys(rs, R, K, True, True, True)
    got = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, 1, A, Ainv)
    o = ko.sort_indices(ko.key_compile(ko.monomial_key(A.perm, A.scale), ko.toeplitz_conv2d((C, U, V), f, b, 1), ko.monomial_key(Ainv.perm, Ainv.scale)))
    _same(got.csr_arrays(), (o.indptr, o.indices, o.data), 'oracle')


@pytest.mark.parametrize('with_last', [False, True])
def test_fused_linear_row_shard_is_one_group(with_last):
    """A row shard of a linear layer (dist.py) keeps every row in ONE pattern group even when a weight is exactly zero
    (a residual dense row on the CSR kernel costs milliseconds at VGG16 fc6 size)."""
    from keynet_b200 import sparse
    rs = np.random.RandomState(8)
    (n_out, n_in) = (200, 5000)
    w = rs.randn(n_out, n_in).astype(np.float32); w[3, 7] = 0; w[150, 4999] = 0
    b = rs.randn(n_out).astype(np.float32)
    (A, Ainv) = _keys(rs, n_out + 1, n_in + 1, True, True, True)
    rows = rs.permutation(n_out)[:77].astype(np.int64)
    if with_last:
        rows = np.concatenate([rows[:30], [n_out], rows[30:]])
    n_phys = n_in + 1 + 5
    remap = np.concatenate([rs.permutation(n_phys - 1)[:n_in], [n_phys - 1]]).astype(np.int64)
    got = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv, rows=rows, col_remap=remap, n_cols_phys=n_phys)
    try:
        sparse.direct_compile_enabled(False)
        ref = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv, rows=rows, col_remap=remap, n_cols_phys=n_phys)
    finally:
        sparse.direct_compile_enabled(True)
    _same(got.csr_arrays(), ref.csr_arrays(), 'sharded linear CSR')
    assert got._pg is not None and len(got._pg.classes) == 1 and got._pg.classes[0]['G'] == 77
    assert (got._pg.rest is None) == (not with_last)
    _same(_groups_to_csr(got), ref.csr_arrays(), 'sharded linear groups')
    only = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv, rows=rows, col_remap=remap, n_cols_phys=n_phys, want_csr=False)
    assert only._data is None and only.nnz() == ref.nnz()
    X = torch.randn(n_phys, 64, device='cuda')
    y = sparse.spmm(only, X)
    y_ref = sparse.spmm(sparse.SparseMatrix((ref.shape, *ref.csr_arrays())), X)
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-5 * float(y_ref.abs().max()))


def test_fused_linear_wide_rows_sorted_in_global_memory():
    """More columns than the shared-memory sort holds (VGG16 fc6: 25 089): the single group's column list is sorted once in a
    scratch buffer and shared by the whole grid."""
    from keynet_b200 import sparse
    rs = np.random.RandomState(9)
    (n_out, n_in) = (70, 9001)
    w = rs.randn(n_out, n_in).astype(np.float32); w[5, 100] = 0
    b = rs.randn(n_out).astype(np.float32)
    for gain in (False, True):
        (A, Ainv) = _keys(rs, n_out + 1, n_in + 1, True, True, gain)
        got = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv)
        try:
            sparse.direct_compile_enabled(False)
            ref = sparse.keyed_linear(torch.from_numpy(w), torch.from_numpy(b), A, Ainv)
        finally:
            sparse.direct_compile_enabled(True)
        _same(got.csr_arrays(), ref.csr_arrays(), 'wide linear CSR')
