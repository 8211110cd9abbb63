"""Host-side key logic (CPU): keygen / block permutation / key algebra against reference goldens.
The keys are integer permutations and fp32 gains; parity is bit-exact."""
import json

import numpy as np
import pytest

from keynet_b200 import system, sparse, blockpermute, util
from tests import golden_util as gu


def _names(z):
    return gu.jstr(z, 'names')


@pytest.mark.parametrize('name', ['perm', 'gain', 'perm_gain', 'hier', 'hier_small', 'hrot', 'blockorder_perm', 'local_perm', 'local_gain', 'fc_perm'])
def test_keygen_matches_reference(name):
    z = gu.load('keygen_kat.npz')
    kw = gu.jstr(z, name + '.args')
    shape = tuple(kw.pop('shape'))
    for k in ('hierarchical_blockshape', 'hierarchical_permute_at_level'):
        if k in kw:
            kw[k] = tuple(kw[k])
    np.random.seed(11)
    (A, Ainv) = system.keygen(shape, **kw)
    for (K, pre) in ((A, name + '.A'), (Ainv, name + '.Ainv')):
        (perm, scale) = gu.monomial_from_coo(z, pre)
        assert np.array_equal(K.perm, perm), pre
        assert np.array_equal(K.scale.view(np.uint32), scale.view(np.uint32)), pre
    # A . Ainv is the identity permutation with gains within one ulp of 1
    I = A.dot(Ainv)
    assert I.is_unpermuted() and np.allclose(I.scale, 1.0, atol=2e-7)


@pytest.mark.parametrize('name', ['bias', 'affine'])
def test_keygen_bias_keys_match_reference(name):
    """Affine photometric keys [[D P, b],[0, 1]] and their Woodbury inverses: dense forms bit-equal to the reference."""
    import scipy.sparse
    z = gu.load('keygen_kat.npz')
    kw = gu.jstr(z, name + '.args')
    shape = tuple(kw.pop('shape'))
    np.random.seed(11)
    (A, Ainv) = system.keygen(shape, **kw)
    assert A.has_bias() and Ainv.has_bias()
    for (K, pre) in ((A, name + '.A'), (Ainv, name + '.Ainv')):
        (shp, row, col, data) = gu.coo_arrays(z, pre)
        ref = np.asarray(scipy.sparse.coo_matrix((data, (row, col)), shape=shp).todense(), dtype=np.float32)
        assert np.array_equal(K.todense().view(np.uint32), ref.view(np.uint32)), pre
    I = A.dot(Ainv).todense()
    assert np.allclose(I, np.eye(I.shape[0]), atol=1e-5)


@pytest.mark.parametrize('name', ['givens_local', 'givens_global', 'doubly_stochastic', 'orthogonal_tiled', 'givens_fc'])
def test_keygen_general_keys_match_reference(name):
    """Givens-orthogonal / doubly-stochastic key families (several entries per row): same RNG draws, same structure, values
    equal to the reference's (bit-equal for the Givens families; the doubly stochastic block passes through a dense
    float64 inverse, compared to 1e-6)."""
    import scipy.sparse
    from keynet_b200.sparse import SparseKey
    z = gu.load('keygen_kat.npz')
    kw = gu.jstr(z, name + '.args')
    shape = tuple(kw.pop('shape'))
    for k in ('hierarchical_blockshape', 'hierarchical_permute_at_level', 'tileshape'):
        if k in kw:
            kw[k] = tuple(kw[k])
    np.random.seed(11)
    (A, Ainv) = system.keygen(shape, **kw)
    tail = np.random.rand()
    np.random.seed(11)
    assert isinstance(A, SparseKey)
    for (K, pre) in ((A, name + '.A'), (Ainv, name + '.Ainv')):
        (shp, row, col, data) = gu.coo_arrays(z, pre)
        ref = np.asarray(scipy.sparse.coo_matrix((data.astype(np.float64), (row, col)), shape=shp).todense())
        got = K.todense().astype(np.float64)
        assert np.array_equal(got != 0, ref != 0), pre
        if name == 'doubly_stochastic':
            assert np.allclose(got, ref, rtol=1e-5, atol=1e-6), pre
        else:
            assert np.array_equal(got.astype(np.float32), ref.astype(np.float32)), pre


def test_keygen_general_geometric_keys_are_inverse_pairs():
    """Givens-orthogonal / doubly stochastic options give general sparse keys (sparse.SparseKey) with A . Ainv = I."""
    from keynet_b200.sparse import SparseKey
    for kw in [dict(global_geometric='givens_orthogonal', alpha=9), dict(local_geometric='doubly_stochastic', alpha=2, blocksize=4),
               dict(local_geometric='givens_orthogonal', alpha=3, blocksize=4, local_photometric='uniform_random_affine', beta=1.0, gamma=1.0)]:
        args = dict(global_geometric='identity', local_geometric='identity', global_photometric='identity', local_photometric='identity')
        args.update(kw)
        np.random.seed(4)
        (A, Ainv) = system.keygen((2, 8, 8), **args)
        assert isinstance(A, SparseKey) and A.shape == (129, 129)
        assert np.allclose(A.todense().astype(np.float64) @ Ainv.todense().astype(np.float64), np.eye(129), atol=1e-4)


def test_keygen_rejects_unknown_options():
    with pytest.raises(ValueError):
        system.keygen((1, 8, 8), 'nope', 'identity', 'identity', 'identity')
    with pytest.raises(ValueError):
        system.keygen((1, 8, 8), 'identity', 'identity', 'identity', 'identity', memoryorder='nope')
    with pytest.raises(AssertionError):     # global permutation is not tile compressible (system.py:360)
        system.keygen((1, 8, 8), 'permutation', 'identity', 'identity', 'identity', tileshape=(4, 4))


def test_hierarchical_block_permutation_matches_reference():
    z = gu.load('blockpermute_kat.npz')
    for name in _names(z):
        kw = gu.jstr(z, name + '.args')
        P = blockpermute.hierarchical_block_permutation_matrix(tuple(kw['imgshape']), tuple(kw['blockshape']), kw['permute_at_level'],
                                                               min_blocksize=8, seed=42, twist=kw['twist'], strict=False)
        assert np.array_equal(P.perm, z[name + '.cols']), name
        assert np.array_equal(np.sort(P.perm), np.arange(len(P.perm))), name


def test_block_permute_image_form_equals_matrix_form():
    """reference test/test_blockpermute.py:62 -- matrix form == image form (seed 42)."""
    rs = np.random.RandomState(0)
    img = rs.randint(0, 255, size=(64, 64, 3)).astype(np.uint8)
    out = blockpermute.hierarchical_block_permute(img, (2, 2), [0, 1], min_blocksize=8, seed=42)
    P = blockpermute.hierarchical_block_permutation_matrix(img.shape, (2, 2), [0, 1], min_blocksize=8, seed=42)
    assert np.array_equal(img.flatten()[P.perm].reshape(img.shape), out)
    with pytest.raises(ValueError):
        blockpermute.hierarchical_block_permute(np.zeros((16, 16, 1)), (2, 2), [0, 1, 2], min_blocksize=8, seed=0)


def test_monomial_key_algebra():
    rs = np.random.RandomState(3)
    n = 50
    A = sparse.MonomialKey(rs.permutation(n), rs.rand(n).astype(np.float32) + 0.5)
    B = sparse.MonomialKey(rs.permutation(n), rs.rand(n).astype(np.float32) + 0.5)
    assert np.allclose(A.dot(B).todense(), A.todense().dot(B.todense()), rtol=1e-6)
    assert np.array_equal(A.transpose().todense(), A.todense().T)
    assert sparse.sparse_identity_matrix(n).is_identity()
    H = sparse.sparse_affine_to_linear(A)
    assert H.shape == (n + 1, n + 1) and H.perm[-1] == n and H.scale[-1] == 1.0
    # products are rounded once in fp32, like scipy's SpGEMM on single-entry rows
    import scipy.sparse
    ref = A.toscipy().dot(B.toscipy()).tocsr(); ref.sort_indices()
    got = A.dot(B).toscipy(); got.sort_indices()
    assert np.array_equal(ref.indices, got.indices) and np.array_equal(ref.data.view(np.uint32), got.data.view(np.uint32))


def test_lazy_identity_and_unscaled_keys_equal_explicit_ones():
    """MonomialKey.identity(n) and keys without a value array (all ones) never build their 3.2 M-element arrays on the VGG16
    keying path; every product / transpose / augmentation must still be bit-identical to the explicit form."""
    rs = np.random.RandomState(4)
    n = 40
    I = sparse.MonomialKey.identity(n)
    Ie = sparse.MonomialKey(np.arange(n), np.ones(n, dtype=np.float32))
    P = sparse.MonomialKey(rs.permutation(n))                                            # unscaled: no value array
    Pe = sparse.MonomialKey(P.perm.copy(), np.ones(n, dtype=np.float32))
    D = sparse.MonomialKey(np.arange(n), (rs.rand(n) + 0.5).astype(np.float32))          # diagonal gain
    G = sparse.MonomialKey(rs.permutation(n), (rs.rand(n) + 0.5).astype(np.float32))
    assert I._perm is None and I._scale is None and I.is_identity() and I.nnz == n and P._scale is None
    assert P.is_unscaled() and not P.is_unpermuted() and D.is_unpermuted() and not D.is_unscaled()

    def same(X, Y):
        return np.array_equal(X.perm, Y.perm) and np.array_equal(X.scale.view(np.uint32), Y.scale.view(np.uint32)) and X.shape == Y.shape

    for (X, Xe) in ((I, Ie), (P, Pe)):
        for K in (P, D, G, I):
            assert same(X.dot(K), Xe.dot(K)) and same(K.dot(X), K.dot(Xe))
        assert same(X.transpose(), Xe.transpose())
        assert same(sparse.sparse_affine_to_linear(X), sparse.sparse_affine_to_linear(Xe))
    assert I.dot(G) is G and G.dot(I) is G and D.transpose() is D                        # no copies for the trivial cases
    assert sparse.sparse_affine_to_linear(I).is_identity() and sparse.sparse_affine_to_linear(I)._perm is None
    assert same(P.dot(P.transpose()), Ie) and P.dot(P.transpose()).is_identity()
    assert np.array_equal(I.todense(), np.eye(n, dtype=np.float32))
    # a homogeneous gain key keeps its trailing one
    H = sparse.sparse_affine_to_linear(D)
    assert H.perm[-1] == n and H.scale[-1] == 1.0 and np.array_equal(H.scale[:n], D.scale)


def test_keyed_model_plan_merges_and_dropout_links():
    """KeyedModel's three passes on the host: dropout leaves the links but keeps its table entry (it still draws a key), a layer
    followed by ReLU / its batchnorm is keyed by the follower with (A_f . A_f_in) . A_layer."""
    import torch
    from collections import OrderedDict
    from torch import nn
    table = OrderedDict([('input', {'prevlayer': None, 'nextlayer': 'fc1'}), ('fc1', {'prevlayer': 'input', 'nextlayer': 'dropout1'}),
                         ('dropout1', {'prevlayer': 'fc1', 'nextlayer': 'dropout2'}), ('dropout2', {'prevlayer': 'dropout1', 'nextlayer': 'fc2'}),
                         ('fc2', {'prevlayer': 'dropout2', 'nextlayer': 'output'}), ('output', {'prevlayer': 'fc2', 'nextlayer': None})])
    t = system._unlink(table, 'dropout')
    assert list(t) == list(table) and t['fc1']['nextlayer'] == 'fc2' and t['fc2']['prevlayer'] == 'fc1'

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(1, 2, 3, padding=1); self.conv1_bn = nn.BatchNorm2d(2); self.conv2 = nn.Conv2d(2, 2, 3, padding=1)
            self.relu2 = nn.ReLU(); self.dropout2 = nn.Dropout(); self.fc = nn.Linear(2 * 4 * 4, 3)

        def forward(self, x):
            return self.fc(self.dropout2(self.relu2(self.conv2(self.conv1_bn(self.conv1(x))))).flatten(1))
    net = Net().eval()
    from keynet_b200 import torch as ktorch
    shapes = system._unlink(ktorch.netshape(net, (1, 4, 4)), 'dropout')
    rs = np.random.RandomState(0)
    keys = {}
    for (k, v) in shapes.items():
        if k in ('input', 'output'):
            continue
        n = int(np.prod(v['outshape'])) + 1
        A = sparse.MonomialKey(np.concatenate([rs.permutation(n - 1), [n - 1]]), np.concatenate([rs.rand(n - 1) + 0.5, [1.0]]).astype(np.float32))
        keys[k] = (A, sparse.MonomialKey(A.transpose().perm, (1.0 / A.transpose().scale).astype(np.float32)))
    inkey = sparse.MonomialKey.identity(17)
    out_key = {k: A for (k, (A, _)) in keys.items()}
    in_key = {k: (inkey if shapes[k]['prevlayer'] == 'input' else keys[shapes[k]['prevlayer']][1]) for k in keys}
    jobs = list(system.KeyedModel._plan(net, shapes, out_key, in_key))
    assert [(j.name, j.relu_name) for j in jobs] == [('conv1', None), ('conv2', 'relu2'), ('fc', None)]
    (c1, c2, fc) = jobs
    assert c1.module is not net.conv1 and not torch.equal(c1.module.weight, net.conv1.weight)      # batchnorm folded into a copy
    expect = out_key['conv1_bn'].dot(in_key['conv1_bn']).dot(out_key['conv1'])
    assert np.array_equal(c1.A.perm, expect.perm) and np.array_equal(c1.A.scale, expect.scale) and c1.Ainv is inkey
    expect = out_key['relu2'].dot(in_key['relu2']).dot(out_key['conv2'])
    assert np.array_equal(c2.A.perm, expect.perm) and np.array_equal(c2.A.scale, expect.scale) and c2.Ainv is keys['conv1_bn'][1]
    assert fc.Ainv is keys['relu2'][1] and fc.A is out_key['fc']          # the dropout between relu2 and fc is gone from the links


def test_permutation_consumes_same_rng_stream_as_reference_form():
    np.random.seed(5)
    a = np.random.permutation(list(range(0, 37)))      # reference form (sparse.py:283)
    np.random.seed(5)
    b = sparse.sparse_permutation_matrix(37).perm
    assert np.array_equal(a, b)


def test_find_closest_positive_divisor_and_blockview():
    assert util.find_closest_positive_divisor(28, 8) == 7
    assert util.find_closest_positive_divisor(224, 56) == 56
    assert util.find_closest_positive_divisor(14, 56) == 14
    assert util.find_closest_positive_divisor(7, 2) == 7
    A = np.arange(36).reshape(6, 6)
    assert np.array_equal(util.blockview(A, 3)[1, 0], A[3:6, 0:3])


def test_mat2gray_key_maps_onto_unit_interval_and_inverts():
    rs = np.random.RandomState(2)
    x = (rs.randn(40) * 7 + 3).astype(np.float32)
    (A, Ainv) = sparse.mat2gray(x)
    xl = np.concatenate([x, [1.0]]).reshape(-1, 1).astype(np.float32)
    y = A.apply(xl)
    assert abs(float(y[:-1].min())) < 1e-6 and abs(float(y[:-1].max()) - 1.0) < 1e-6 and y[-1, 0] == 1.0
    assert np.allclose(Ainv.apply(y), xl, atol=1e-4)
    K = sparse.SparseKey.from_monomial(A)
    assert np.allclose(K.apply(xl), y, atol=1e-6)


def test_spy_renders_keys_and_blocks():
    P = sparse.sparse_permutation_matrix(300)
    im = sparse.spy(P, mindim=64, showdim=128)
    assert im.mode == 'RGB' and max(im.size) == 128
    im2 = sparse.spy(sparse.SparseKey.from_monomial(P), mindim=512, showdim=256, range=(0, 100), eps=0.5)
    assert max(im2.size) == 256


def test_pixel_tile_hint_clusters_rows_by_output_tile():
    """Spatial cluster ids of the clustered kernel (sparse._pixel_tile_hint): rows whose Toeplitz row lies in the same tile of
    output pixels share an id whatever the output key; the homogeneous row gets an id of its own."""
    import torch
    from keynet_b200.sparse import _pixel_tile_hint
    (M, Uo, Vo, C, k) = (6, 28, 28, 1, 3)
    rs = np.random.RandomState(0)
    perm = np.concatenate([rs.permutation(M * Uo * Vo), [M * Uo * Vo]])           # output key: row r of W_hat = Toeplitz row perm[r]
    hint = _pixel_tile_hint(torch.from_numpy(perm), len(perm), (M, Uo, Vo), C, k, 1, False, 'cpu').numpy()
    px = perm[:-1] % (Uo * Vo)
    (py, pxx) = (px // Vo, px % Vo)
    t = 7                                                                            # largest tile whose union fits, dividing 28
    assert np.array_equal(hint[:-1], (py // t) * (Vo // t) + pxx // t)
    assert hint[-1] == (Uo // t) * (Vo // t) and hint[-1] not in hint[:-1]
    # all channels of a pixel land in the same cluster; a cluster holds t*t pixels x M channels
    assert np.all(np.bincount(hint[:-1]) == t * t * M)
    # depthwise (pooling): clusters are per channel
    hp = _pixel_tile_hint(None, M * 14 * 14 + 1, (M, 14, 14), M, 3, 2, True, 'cpu').numpy()
    assert len(np.unique(hp[:-1])) % M == 0 and hp[-1] not in hp[:-1]
    # a reduction too long for the staging buffer: no clustering
    assert _pixel_tile_hint(None, 10, (4, 3, 3), 512, 3, 1, False, 'cpu') is None


def test_key_serialisation_roundtrip_all_key_kinds():
    """io._key_state / _key_load: monomial, monomial + bias column and general sparse keys survive the flat-tensor form."""
    from keynet_b200 import io as kio
    from keynet_b200.sparse import SparseKey, MonomialKey
    np.random.seed(2)
    (A, _) = system.keygen((2, 4, 4), 'permutation', 'identity', 'uniform_random_gain', 'identity', beta=1.0)
    (B, _) = system.keygen((2, 4, 4), 'permutation', 'identity', 'uniform_random_affine', 'identity', beta=1.0, gamma=1.0)
    (C, _) = system.keygen((2, 4, 4), 'identity', 'givens_orthogonal', 'identity', 'identity', alpha=3, blocksize=2)
    for K in (A, B, C):
        K2 = kio._key_load(kio._key_state(K))
        assert type(K2) is type(K)
        assert np.array_equal(K2.todense(), K.todense())
    assert kio._key_load(kio._key_state(None)) is None
    assert isinstance(C, SparseKey) and isinstance(B, MonomialKey) and B.has_bias()
