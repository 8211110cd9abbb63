"""Spatially tiled tensor-core kernel (csrc/pgtile_tc.cu) against the oracle's csr_matvecs on the same compiled CSR: conv layers
with G <= 128 output channels, permutation keys on both sides, stride 1 / 2, 2x2 and 1x2 tiles, image borders, ragged batches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _keys(rs, n_out, n_in, permute):
    from keynet_b200.sparse import MonomialKey
    po = np.concatenate([rs.permutation(n_out - 1) if permute else np.arange(n_out - 1), [n_out - 1]])
    pi = np.concatenate([rs.permutation(n_in - 1) if permute else np.arange(n_in - 1), [n_in - 1]])
    return (MonomialKey(po), MonomialKey(pi))


CASES = [  # (C, U, M, k, stride)
    (16, 8, 32, 3, 1), (32, 6, 64, 3, 1), (64, 4, 96, 3, 1), (128, 4, 128, 3, 1), (16, 8, 64, 3, 2), (48, 6, 80, 3, 1)]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('permute', [False, True])
@pytest.mark.parametrize('N', [128, 200, 384])
def test_tile_kernel_matches_oracle(case, permute, N):
    from keynet_b200 import sparse
    from oracle import keynet_oracle as ko
    (C, U, M, k, stride) = case
    rs = np.random.RandomState(C + M + N)
    f = rs.randn(M, C, k, k).astype(np.float32)
    b = rs.randn(M).astype(np.float32)
    (R, K) = (M * (U // stride) ** 2 + 1, C * U * U + 1)
    (A, Ainv) = _keys(rs, R, K, permute)
    W = sparse.keyed_toeplitz_conv2d((C, U, U), f, b, stride, A, Ainv)
    cls = W._pg.classes[0]
    assert cls.get('tile') is not None, 'layer should qualify for the tiled kernel'
    (ip, ix, dt) = W.csr_arrays()
    X = rs.randn(K, N).astype(np.float32)
    X[-1] = 1.0
    for relu in (False, True):
        ref = ko.spmm(ko.csr(W.shape, ip, ix, dt), X, relu=relu, threads=8)
        y = sparse.spmm(W, torch.from_numpy(X).cuda(), relu=relu).cpu().numpy()
        err = np.abs(y - ref)
        assert np.all(err <= 1e-4 * np.abs(ref) + 1e-5 * np.abs(ref).max()), (case, permute, N, relu, float(err.max()), float(np.abs(ref).max()))
    # the per-pixel kernel gives the same answer (A/B switch)
    try:
        sparse.tiles_enabled(False)
        y2 = sparse.spmm(W, torch.from_numpy(X).cuda(), relu=True).cpu().numpy()
    finally:
        sparse.tiles_enabled(True)
    assert np.allclose(y, y2, rtol=1e-4, atol=1e-5 * np.abs(ref).max())


def test_tile_kernel_not_used_with_gain_keys_or_odd_shapes():
    from keynet_b200 import sparse
    from keynet_b200.sparse import MonomialKey
    rs = np.random.RandomState(0)
    f = rs.randn(32, 16, 3, 3).astype(np.float32); b = rs.randn(32).astype(np.float32)
    (R, K) = (32 * 64 + 1, 16 * 64 + 1)
    A = MonomialKey(np.arange(R), np.concatenate([rs.rand(R - 1) + 0.5, [1.0]]).astype(np.float32))
    W = sparse.keyed_toeplitz_conv2d((16, 8, 8), f, b, 1, A, MonomialKey(np.arange(K)))
    assert W._pg.classes[0].get('tile') is None                   # gains make every pixel's weights its own
    f = rs.randn(32, 16, 3, 3).astype(np.float32)
    W = sparse.keyed_toeplitz_conv2d((16, 7, 7), f, b, 1, None, MonomialKey(np.arange(16 * 49 + 1)))
    assert W._pg.classes[0].get('tile') is None                   # 7 x 7 outputs do not divide into 2 x 2 tiles
