"""Pin the CPU oracle (oracle/) against the reference's golden vectors and against scipy's kernels.

CPU-only (-m "not gpu").  Until these pass, the oracle cannot be trusted as the checker for
the CUDA path.  Goldens come from the UNMODIFIED reference (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from oracle import keynet_oracle as ko
from tests import golden_util as gu


def _assert_csr_equal(A, shape, indptr, indices, data, what=''):
    assert tuple(A.shape) == tuple(shape), what
    assert np.array_equal(A.indptr, indptr), what + ' indptr'
    assert np.array_equal(A.indices, indices), what + ' indices'
    # bit-exact values (compare the raw fp32 bit patterns, so -0.0 / NaN cannot hide)
    assert np.array_equal(A.data.view(np.uint32), np.asarray(data, dtype=np.float32).view(np.uint32)), what + ' data bits'


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['conv_s2', 'conv_s1', 'conv_k1', 'conv_k5', 'conv_tiny'])
def test_toeplitz_conv2d_matches_reference(name):
    z = gu.load('toeplitz_kat.npz')
    W = ko.toeplitz_conv2d(tuple(z[name + '.inshape']), z[name + '.f'], bias=z[name + '.b'], stride=int(z[name + '.stride']))
    _assert_csr_equal(W, *gu.csr_arrays(z, name + '.W'), what=name)
    # functional: W . x == conv2d(x)  (reference test/test_sparse.py:223, atol 1e-5)
    x = z[name + '.x']
    y = ko.keyed_forward([(W, False)], ko.affine_to_linear(x))
    assert np.allclose(ko.linear_to_affine(y).reshape(z[name + '.y'].shape), z[name + '.y'], atol=1e-5)


@pytest.mark.parametrize('name', ['pool_s2', 'pool_c3', 'pool_s1'])
def test_toeplitz_avgpool2d_matches_reference(name):
    z = gu.load('toeplitz_kat.npz')
    W = ko.toeplitz_avgpool2d(tuple(z[name + '.inshape']), int(z[name + '.k']), int(z[name + '.stride']))
    _assert_csr_equal(W, *gu.csr_arrays(z, name + '.W'), what=name)


def test_toeplitz_keeps_explicit_zeros_and_collapses_tiny_weights():
    z = gu.load('toeplitz_kat.npz')
    W = ko.toeplitz_conv2d(tuple(z['conv_s1.inshape']), z['conv_s1.f'], bias=z['conv_s1.b'], stride=1)
    assert (W.data == 0).sum() > 0                      # stored zeros (offset trick, sparse.py:184-187)
    assert not np.any(np.abs(W.data[W.data != 0]) < 1e-8)   # 1e-9 became exactly 0


# ---------------------------------------------------------------------------------------------
def _rand_csr(rs, R, C, density, with_zeros=True):
    import scipy.sparse
    A = scipy.sparse.random(R, C, density=density, format='csr', dtype=np.float32, random_state=rs)
    if with_zeros and A.nnz > 3:
        A.data[:: 7] = 0.0
    return A


def test_matmat_bit_exact_vs_scipy_including_stored_order():
    rs = np.random.RandomState(0)
    for (R, K, C, d) in [(50, 40, 60, 0.1), (200, 300, 100, 0.02), (7, 3, 5, 0.9), (30, 30, 30, 0.0)]:
        A = _rand_csr(rs, R, K, d); B = _rand_csr(rs, K, C, d)
        ref = A.dot(B)
        out = ko.matmat(ko.csr(A.shape, A.indptr, A.indices, A.data), ko.csr(B.shape, B.indptr, B.indices, B.data))
        _assert_csr_equal(out, ref.shape, ref.indptr, ref.indices, ref.data, what='matmat %s' % str((R, K, C)))


def test_spmm_bit_exact_vs_scipy():
    rs = np.random.RandomState(1)
    for (R, C, N, d) in [(64, 48, 1, 0.2), (64, 48, 5, 0.2), (300, 500, 33, 0.05), (10, 10, 4, 0.0)]:
        A = _rand_csr(rs, R, C, d, with_zeros=False)
        # unsorted stored order, like the reference's keyed matrices
        for i in range(R):
            s = slice(A.indptr[i], A.indptr[i + 1]); p = rs.permutation(A.indptr[i + 1] - A.indptr[i])
            A.indices[s] = A.indices[s][p]; A.data[s] = A.data[s][p]
        X = rs.randn(C, N).astype(np.float32)
        ref = np.asarray(A.dot(np.matrix(X)))
        for threads in (1, 4):
            out = ko.spmm(ko.csr(A.shape, A.indptr, A.indices, A.data), X, threads=threads)
            assert np.array_equal(out.view(np.uint32), ref.astype(np.float32).view(np.uint32))


def test_coo_tocsr_sums_duplicates_like_scipy():
    import scipy.sparse
    rs = np.random.RandomState(2)
    row = rs.randint(0, 20, 300).astype(np.int32); col = rs.randint(0, 15, 300).astype(np.int32); val = rs.randn(300).astype(np.float32)
    ref = scipy.sparse.coo_matrix((val, (row, col)), shape=(20, 15)).tocsr()
    out = ko.csr_from_coo((20, 15), row, col, val)
    # scipy sorts rows with an unstable std::sort, so the order in which duplicates are summed is
    # unspecified: structure must match exactly, merged values to fp32 rounding.  (The keyed path
    # never produces duplicate (row, col) pairs -- Toeplitz triplets are unique.)
    assert np.array_equal(out.indptr, ref.indptr) and np.array_equal(out.indices, ref.indices)
    assert np.allclose(out.data, ref.data, rtol=1e-5, atol=1e-6)
    # without duplicates the result is bit-exact
    (u, idx) = np.unique(row.astype(np.int64) * 15 + col, return_index=True)
    ref = scipy.sparse.coo_matrix((val[idx], (row[idx], col[idx])), shape=(20, 15)).tocsr()
    out = ko.csr_from_coo((20, 15), row[idx], col[idx], val[idx])
    _assert_csr_equal(out, ref.shape, ref.indptr, ref.indices, ref.data)


# ---------------------------------------------------------------------------------------------
def _lenet_layers_from_golden(z):
    """Rebuild every keyed LeNet layer with the oracle from weights + the recorded reference keys.

    Key chaining follows keynet/system.py:43-51,85-92 (conv->relu merge: A = (A_relu.A_conv^-1).A_conv)."""
    keys = {}
    order = ['input', 'conv1', 'relu1', 'pool1', 'conv2', 'relu2', 'pool2', 'fc1', 'relu3', 'fc2', 'relu4', 'fc3']
    assert int(z['n_keygen_calls']) == len(order)
    for (i, name) in enumerate(order):
        keys[name] = (ko.csr_from_coo(*gu.coo_arrays(z, 'keygen.%d.A' % i)), ko.csr_from_coo(*gu.coo_arrays(z, 'keygen.%d.Ainv' % i)))

    def merged(relu, prev):
        B = ko.matmat(keys[relu][0], keys[prev][1])
        return ko.matmat(B, keys[prev][0])
    w = lambda k: z['weights.' + k]
    W = {}
    W['conv1'] = ko.key_compile(merged('relu1', 'conv1'), ko.toeplitz_conv2d((1, 28, 28), w('conv1.weight'), w('conv1.bias'), 1), keys['input'][1])
    W['pool1'] = ko.key_compile(keys['pool1'][0], ko.toeplitz_avgpool2d((6, 28, 28), 3, 2), keys['relu1'][1])
    W['conv2'] = ko.key_compile(merged('relu2', 'conv2'), ko.toeplitz_conv2d((6, 14, 14), w('conv2.weight'), w('conv2.bias'), 1), keys['pool1'][1])
    W['pool2'] = ko.key_compile(keys['pool2'][0], ko.toeplitz_avgpool2d((16, 14, 14), 3, 2), keys['relu2'][1])
    W['fc1'] = ko.key_compile(merged('relu3', 'fc1'), ko.linear_matrix(w('fc1.weight'), w('fc1.bias')), keys['pool2'][1])
    W['fc2'] = ko.key_compile(merged('relu4', 'fc2'), ko.linear_matrix(w('fc2.weight'), w('fc2.bias')), keys['relu3'][1])
    W['fc3'] = ko.key_compile(None, ko.linear_matrix(w('fc3.weight'), w('fc3.bias')), keys['relu4'][1])
    relu_after = dict(conv1=True, pool1=False, conv2=True, pool2=False, fc1=True, fc2=True, fc3=False)
    return (W, relu_after)


@pytest.mark.parametrize('golden', ['lenet_cfg1.npz', 'lenet_cfg3.npz'])
def test_lenet_key_compile_and_forward_bit_exact(golden):
    z = gu.load(golden)
    (W, relu_after) = _lenet_layers_from_golden(z)
    layers = gu.jstr(z, 'layers')
    assert layers == ['conv1', 'pool1', 'conv2', 'pool2', 'fc1', 'fc2', 'fc3']
    for k in layers:
        # stored (unsorted) order, exactly what scipy's SpGEMM leaves behind
        assert np.array_equal(W[k].indices, z['layer.%s.W.stored_indices' % k]), k
        assert np.array_equal(W[k].data.view(np.uint32), z['layer.%s.W.stored_data' % k].view(np.uint32)), k
        _assert_csr_equal(ko.sort_indices(W[k]), *gu.csr_arrays(z, 'layer.%s.W' % k), what=k)
    # forward: bit-exact per layer against the reference's activations
    X = np.ascontiguousarray(z['x_cipher'].T)
    for k in layers:
        Y = ko.spmm(W[k], X)
        assert np.array_equal(np.ascontiguousarray(Y.T).view(np.uint32), z['layer.%s.y' % k].view(np.uint32)), k
        X = np.maximum(Y, 0) if relu_after[k] else Y
    assert np.array_equal(ko.linear_to_affine(np.ascontiguousarray(X.T)), z['logits_keyed'])
    assert np.allclose(z['logits_keyed'], z['logits_plain'], atol=1e-5)


def test_lenet_cfg1_structure_matches_notebook():
    """demo/lenet.ipynb cell 2: per-layer (shape, nnz) of the LeNet PermutationKeynet (RNG independent)."""
    z = gu.load('lenet_cfg1.npz')
    expect = dict(conv1=((4705, 785), 45049), pool1=((1177, 4705), 10087), conv2=((3137, 1177), 156737), pool2=((785, 3137), 6401),
                  fc1=((121, 785), 94201), fc2=((85, 121), 10165), fc3=((11, 85), 851))
    (W, _) = _lenet_layers_from_golden(z)
    for (k, (shape, nnz)) in expect.items():
        assert W[k].shape == shape and len(W[k].data) == nnz
    assert int(z['num_parameters']) == sum(v[1] for v in expect.values())


def test_challenge_known_answer():
    """demo/challenge.ipynb cell 5: fixed matrices, fixed image, printed encoding (4 d.p.)."""
    z = gu.load('challenge_kat.npz')
    layers = gu.jstr(z, 'layers')
    relu_after = dict(conv1=True, pool1=False, conv2=True, pool2=False, fc1=True, fc2=True, fc3=False)
    mats = [(ko.csr(*gu.csr_arrays(z, 'layer.%s.W' % k)), relu_after[k]) for k in layers]
    y = ko.keyed_forward(mats, z['x_linear'])
    enc = ko.linear_to_affine(y).flatten()
    assert np.allclose(enc, z['y_printed'], atol=1e-4)
    assert np.allclose(y, z['y_reference'], rtol=1e-4, atol=1e-6)


def test_linear_to_affine_raises_on_bad_homogeneous_coordinate():
    with pytest.raises(ValueError):
        ko.linear_to_affine(np.array([[1.0, 2.0, 1.1]]))
    assert ko.linear_to_affine(np.array([[1.0, 2.0, 1.0005]])).shape == (1, 2)
