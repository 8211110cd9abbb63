"""General (non-monomial) keys on the host: Givens-orthogonal / doubly stochastic key generation with the reference's
RNG stream, SparseKey algebra, and the oracle's csr_matmat compile of a LeNet keyed with them -- all against a fixture
made by running the unmodified reference (tests/golden/make_golden.py lenet_givens = test/test_keynet.py:180-197)."""
import json

import numpy as np
import pytest
import torch
from torch import nn

from tests import golden_util as gu
from keynet_b200 import system, nets
from keynet_b200.sparse import SparseKey, MonomialKey, sparse_orthogonal_matrix, sparse_block_diagonal_repeat, sparse_affine_to_linear
from oracle import keynet_oracle as ko

CFG = dict(global_geometric='hierarchical_rotation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0),
           global_photometric='uniform_random_bias', local_geometric='givens_orthogonal', alpha=2.0, blocksize=8,
           local_photometric='uniform_random_affine', beta=1.0, gamma=1.0, memoryorder='block')


def _dense_from_coo(z, prefix):
    (shape, row, col, data) = gu.coo_arrays(z, prefix)
    D = np.zeros(shape, dtype=np.float64)
    np.add.at(D, (row, col), data)
    return D


def _golden_net(z):
    net = nets.LeNet_AvgPool().eval()
    net.load_state_dict({k[len('weights.'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('weights.')})
    return net


def test_sparse_key_algebra_matches_dense():
    rs = np.random.RandomState(0)
    D1 = rs.randn(9, 9) * (rs.rand(9, 9) < 0.3); D2 = rs.randn(9, 9) * (rs.rand(9, 9) < 0.3)
    (A, B) = (SparseKey.from_dense(D1.astype(np.float32)), SparseKey.from_dense(D2.astype(np.float32)))
    assert np.allclose(A.dot(B).todense(), D1.astype(np.float32) @ D2.astype(np.float32), atol=1e-6)
    assert np.array_equal(A.transpose().todense(), D1.astype(np.float32).T)
    M = MonomialKey(rs.permutation(9), rs.rand(9).astype(np.float32) + 0.5)
    assert np.allclose(M.dot(A).todense(), M.todense() @ A.todense(), atol=1e-6)
    assert np.allclose(A.dot(M).todense(), A.todense() @ M.todense(), atol=1e-6)
    # exact zeros produced by cancellation are dropped, like scipy's csr_matmat
    X = SparseKey.from_dense(np.array([[1, 1], [0, 0]], dtype=np.float32)); Y = SparseKey.from_dense(np.array([[1, 2], [-1, 3]], dtype=np.float32))
    P = X.dot(Y)
    assert P.nnz == 1 and P.todense()[0, 1] == 5
    # block-diagonal repeat with a ragged identity tail (keynet/sparse.py:657-687), homogeneous augmentation
    R = sparse_block_diagonal_repeat(SparseKey.from_dense(np.arange(1, 10, dtype=np.float32).reshape(3, 3)), (8, 8)).todense()
    assert np.array_equal(R[3:6, 3:6], np.arange(1, 10).reshape(3, 3)) and np.array_equal(R[6:, 6:], np.eye(2)) and R[:3, 3:].sum() == 0
    L = sparse_affine_to_linear(SparseKey.from_dense(np.eye(3, dtype=np.float32) * 2), bias=np.array([1, 0, 3])).todense()
    assert np.array_equal(L, np.array([[2, 0, 0, 1], [0, 2, 0, 0], [0, 0, 2, 3], [0, 0, 0, 1]], dtype=np.float32))


def test_givens_matrix_is_orthogonal_and_sparse():
    np.random.seed(5)
    (S, Sinv) = sparse_orthogonal_matrix(16, 12, withinverse=True)
    assert np.allclose(S.todense() @ Sinv.todense(), np.eye(16), atol=1e-6)
    assert S.nnz < 16 * 16 and S.data.dtype == np.float32


def test_general_sensor_key_matches_reference_bit_for_bit():
    z = gu.load('lenet_givens.npz')
    np.random.seed(0)
    (A, Ainv) = system.keypair_policy(**CFG)('input', (1, 28, 28))
    assert isinstance(A, SparseKey) and isinstance(Ainv, SparseKey)
    assert np.array_equal(A.todense().astype(np.float64), _dense_from_coo(z, 'sensor.A'))
    assert np.array_equal(Ainv.todense().astype(np.float64), _dense_from_coo(z, 'sensor.Ainv'))


def test_oracle_compile_with_general_keys_matches_reference():
    """Host keys (this repo) + csr_matmat restatement (oracle) reproduce every compiled layer of the reference."""
    z = gu.load('lenet_givens.npz')
    net = _golden_net(z)
    got = {}

    def k(K):
        if K is None:
            return None
        K = SparseKey.coerce(K)
        return ko.csr(K.shape, K.indptr, K.indices, K.data.astype(np.float32))

    def f_layergen(module, inshape, outshape, A, Ainv):
        if isinstance(module, nn.Conv2d):
            W = ko.toeplitz_conv2d(inshape, module.weight.detach().numpy(), module.bias.detach().numpy(), module.stride[0])
        elif isinstance(module, nn.AvgPool2d):
            W = ko.toeplitz_avgpool2d(inshape, module.kernel_size, module.stride)
        else:
            W = ko.linear_matrix(module.weight.detach().numpy(), module.bias.detach().numpy())

        class Rec(nn.Module):
            def fuse_relu(self, flag=True):
                return self
        r = Rec(); r.W = ko.sort_indices(ko.key_compile(k(A), W, k(Ainv)))
        return r
    np.random.seed(0)
    f_keypair = system.keypair_policy(**CFG)
    (A, Ainv) = f_keypair('input', (1, 28, 28))
    model = system.KeyedModel(net, (1, 28, 28), Ainv, f_keypair, f_layergen)
    names = gu.jstr(z, 'layers')
    recs = [m for (n, m) in model._keynet.named_children() if hasattr(m, 'W')]
    assert len(recs) == len(names)
    for (name, r) in zip(names, recs):
        (shape, indptr, indices, data) = gu.csr_arrays(z, 'layer.%s.W' % name)
        ref = ko.sort_indices(ko.csr(shape, indptr, indices, data.astype(np.float32)))
        assert r.W.shape == tuple(shape), name
        assert np.array_equal(r.W.indptr, ref.indptr) and np.array_equal(r.W.indices, ref.indices), name
        # identical except a few bias-column entries (long cancelling sums of the affine keys: |diff| < 1e-6)
        assert np.allclose(r.W.data, ref.data, rtol=1e-5, atol=2e-6), name
        assert np.mean(r.W.data == ref.data) > 0.95, name
