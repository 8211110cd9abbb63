"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference goldens.

All tests here need a B200 (`-m gpu`).  Bars: bit-exact for index arrays and compiled values
(permutation / gain keys); fp32 outputs within rtol 1e-4 (north star) -- the summation order of a
warp-shuffle / per-lane accumulation differs from scipy's sequential loop, nothing else does.
"""
import numpy as np
import pytest
import torch

from tests import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _ko():
    from oracle import keynet_oracle as ko
    return ko


def _assert_bit_exact(W, shape, indptr, indices, data, what=''):
    (ip, ix, dt) = W.csr_arrays()
    assert tuple(W.shape) == tuple(shape), what
    assert np.array_equal(ip - ip[0], np.asarray(indptr) - indptr[0]), what + ' indptr'
    assert np.array_equal(ix, indices), what + ' indices'
    assert np.array_equal(dt.view(np.uint32), np.asarray(data, dtype=np.float32).view(np.uint32)), what + ' data bits'


def _close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    atol = 1e-5 * (float(np.abs(b).max()) if b.size else 1.0)          # rtol 1e-4, atol 1e-5 * max|y| (no floor)
    return np.allclose(a, b, rtol=rtol, atol=atol)


# ---------------------------------------------------------------------------------------------
# SpMM
def _rand_csr(rs, R, C, density, empty_rows=True):
    ko = _ko()
    nnz_per_row = rs.binomial(C, density, size=R)
    if empty_rows and R > 3:
        nnz_per_row[rs.randint(0, R, size=max(1, R // 10))] = 0
    indptr = np.concatenate([[0], np.cumsum(nnz_per_row)]).astype(np.int64)
    indices = np.concatenate([np.sort(rs.choice(C, n, replace=False)) for n in nnz_per_row] + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    data = rs.randn(len(indices)).astype(np.float32)
    return ko.csr((R, C), indptr, indices, data)


@pytest.mark.parametrize('N', [1, 2, 3, 4, 5, 7, 8, 9, 17, 31, 32, 33, 64, 96, 100, 128, 130, 256, 1000, 4096])
@pytest.mark.parametrize('relu', [False, True])
def test_spmm_matches_oracle(N, relu):
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(N)
    for (R, C, d) in [(257, 300, 0.1), (64, 1000, 0.3), (5, 7, 0.9), (1000, 50, 0.02)]:
        A = _rand_csr(rs, R, C, d)
        X = rs.randn(C, N).astype(np.float32)
        ref = ko.spmm(A, X, relu=relu, threads=4)
        W = sparse.SparseMatrix((A.shape, A.indptr, A.indices, A.data))
        y = sparse.spmm(W, torch.from_numpy(X).cuda(), relu=relu).cpu().numpy()
        assert y.shape == ref.shape
        assert _close(y, ref), (N, R, C, np.abs(y - ref).max())
        if relu:
            assert (y >= 0).all()


def test_spmm_edge_cases():
    from keynet_b200 import sparse, _native
    ko = _ko()
    # empty matrix (no stored entries) -> zeros
    W = sparse.SparseMatrix(((6, 4), np.zeros(7, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float32)))
    y = sparse.spmm(W, torch.ones(4, 40, device='cuda'))
    assert y.shape == (6, 40) and float(y.abs().max()) == 0.0
    # torchdot: host tensor in -> host tensor out, transposed (non-contiguous) view accepted, dtype coerced
    rs = np.random.RandomState(0)
    A = _rand_csr(rs, 33, 20, 0.3)
    W = sparse.SparseMatrix((A.shape, A.indptr, A.indices, A.data))
    xb = rs.randn(9, 20)                       # float64, batch-major
    y = W.torchdot(torch.from_numpy(xb).t())
    assert (not y.is_cuda) and y.dtype == torch.float32 and y.shape == (33, 9)
    assert _close(y.numpy(), ko.spmm(A, xb.T.astype(np.float32)))
    # shape mismatch is an assertion, as in TiledMatrix.torchdot (sparse.py:605)
    with pytest.raises(AssertionError):
        W.torchdot(torch.zeros(21, 3))
    # aliasing X and Y is rejected by the ABI
    x = torch.zeros(20, 16, device='cuda')
    rc = _native.lib().kn_spmm_csr_f32(_native.ptr(W._indptr), _native.ptr(W._indices), _native.ptr(W._data), 20, 20,
                                       _native.ptr(x), 16, _native.ptr(x), 16, 16, 0, None, None)
    assert rc == -1


def test_spmm_homogeneous_coordinate_stays_exactly_one():
    """Last row of every keyed matrix is e_last, so the trailing activation must stay exactly 1.0."""
    from keynet_b200 import sparse
    W = sparse.keyed_toeplitz_conv2d((2, 6, 6), np.random.RandomState(0).randn(3, 2, 3, 3).astype(np.float32), np.ones(3, dtype=np.float32), 1, None,
                                     sparse.sparse_identity_matrix(2 * 36 + 1))
    X = torch.randn(73, 50, device='cuda'); X[-1] = 1.0
    y = sparse.spmm(W, X, relu=True)
    assert bool((y[-1] == 1.0).all())


def test_exclusive_scan():
    from keynet_b200 import _native
    for n in [0, 1, 5, 2047, 2048, 2049, 100000, 3211265]:
        v = torch.randint(0, 5000, (n,), dtype=torch.int64, device='cuda')
        out = torch.empty(n + 1, dtype=torch.int64, device='cuda')
        _native.check(_native.lib().kn_exclusive_scan_i64(_native.ptr(v) if n else None, _native.ptr(out), n, _native.stream_ptr()))
        ref = torch.cat([torch.zeros(1, dtype=torch.int64, device='cuda'), torch.cumsum(v, 0)])
        assert torch.equal(out, ref), n


# ---------------------------------------------------------------------------------------------
# Toeplitz construction
@pytest.mark.parametrize('name', ['conv_s2', 'conv_s1', 'conv_k1', 'conv_k5', 'conv_tiny'])
def test_toeplitz_conv2d_bit_exact_vs_reference(name):
    from keynet_b200 import sparse
    z = gu.load('toeplitz_kat.npz')
    W = sparse.sparse_toeplitz_conv2d(tuple(int(s) for s in z[name + '.inshape']), z[name + '.f'], bias=z[name + '.b'], stride=int(z[name + '.stride']))
    _assert_bit_exact(W, *gu.csr_arrays(z, name + '.W'), what=name)
    # functional (reference test/test_sparse.py:223): W . x == conv2d(x), atol 1e-5
    x = torch.from_numpy(z[name + '.x'])
    from keynet_b200.torch import affine_to_linear, linear_to_affine
    y = W.torchdot(affine_to_linear(x).t()).t()
    assert np.allclose(linear_to_affine(y).numpy().reshape(z[name + '.y'].shape), z[name + '.y'], atol=1e-5)


@pytest.mark.parametrize('name', ['pool_s2', 'pool_c3', 'pool_s1'])
def test_toeplitz_avgpool2d_bit_exact_vs_reference(name):
    from keynet_b200 import sparse
    z = gu.load('toeplitz_kat.npz')
    C = int(z[name + '.inshape'][0]); k = int(z[name + '.k'])
    W = sparse.sparse_toeplitz_avgpool2d(tuple(int(s) for s in z[name + '.inshape']), (C, C, k, k), int(z[name + '.stride']))
    _assert_bit_exact(W, *gu.csr_arrays(z, name + '.W'), what=name)


def test_toeplitz_and_keycompile_vs_oracle_random_shapes():
    """Random conv / pool / linear layers with random permutation + gain keys on both sides: the GPU compile
    must equal the oracle's two SpGEMMs bit-for-bit in canonical form."""
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(123)
    for trial in range(12):
        C = int(rs.randint(1, 5)); M = int(rs.randint(1, 6)); k = int(rs.choice([1, 3, 5])); stride = int(rs.choice([1, 2]))
        U = int(rs.randint(1, 6)) * stride * 2; V = int(rs.randint(1, 6)) * stride * 2
        f = rs.randn(M, C, k, k).astype(np.float32); f[rs.rand(*f.shape) < 0.1] = 0.0
        b = rs.randn(M).astype(np.float32)
        (Rr, Kk) = (M * (U // stride) * (V // stride) + 1, C * U * V + 1)

        def key(n, permute, gain):
            perm = np.concatenate([rs.permutation(n - 1), [n - 1]]) if permute else np.arange(n)
            scale = np.concatenate([rs.rand(n - 1) + 0.5, [1.0]]).astype(np.float32) if gain else np.ones(n, dtype=np.float32)
            return sparse.MonomialKey(perm, scale)
        A = key(Rr, trial % 2 == 0, trial % 3 == 0) if trial % 5 != 4 else None
        Ain = key(Kk, trial % 2 == 1 or trial % 4 == 0, trial % 3 != 1)
        Ainv = Ain.transpose()
        W = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, stride, A, Ainv)
        ref = ko.sort_indices(ko.key_compile(None if A is None else ko.monomial_key(A.perm, A.scale), ko.toeplitz_conv2d((C, U, V), f, b, stride),
                                             ko.monomial_key(Ainv.perm, Ainv.scale)))
        _assert_bit_exact(W, ref.shape, ref.indptr, ref.indices, ref.data, what='conv trial %d' % trial)
        # row shard: rows [r0, r1) equal the same rows of the full compile
        (r0, r1) = (Rr // 3, Rr - Rr // 4)
        Ws = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, stride, A, Ainv, rows=(r0, r1))
        s = slice(ref.indptr[r0], ref.indptr[r1])
        _assert_bit_exact(Ws, (r1 - r0, Kk), ref.indptr[r0:r1 + 1], ref.indices[s], ref.data[s], what='conv shard %d' % trial)
        # avgpool on the output of that conv
        if (U // stride) % 2 == 0 and (V // stride) % 2 == 0:
            pin = (M, U // stride, V // stride)
            (Rp, Kp) = (M * (pin[1] // 2) * (pin[2] // 2) + 1, M * pin[1] * pin[2] + 1)
            Ap = key(Rp, True, trial % 2 == 0); Apin = key(Kp, trial % 2 == 0, True).transpose()
            Wp = sparse.keyed_toeplitz_avgpool2d(pin, 3, 2, Ap, Apin)
            refp = ko.sort_indices(ko.key_compile(ko.monomial_key(Ap.perm, Ap.scale), ko.toeplitz_avgpool2d(pin, 3, 2), ko.monomial_key(Apin.perm, Apin.scale)))
            _assert_bit_exact(Wp, refp.shape, refp.indptr, refp.indices, refp.data, what='pool trial %d' % trial)
        # linear
        (n_out, n_in) = (int(rs.randint(1, 40)), int(rs.randint(1, 300)))
        Wl = rs.randn(n_out, n_in).astype(np.float32); Wl[rs.rand(n_out, n_in) < 0.2] = 0.0
        bl = rs.randn(n_out).astype(np.float32); bl[0] = 0.0
        Al = key(n_out + 1, True, True) if trial % 2 == 0 else None
        Alin = key(n_in + 1, True, trial % 2 == 1).transpose()
        WL = sparse.keyed_linear(torch.from_numpy(Wl), torch.from_numpy(bl), Al, Alin)
        refl = ko.sort_indices(ko.key_compile(None if Al is None else ko.monomial_key(Al.perm, Al.scale), ko.linear_matrix(Wl, bl), ko.monomial_key(Alin.perm, Alin.scale)))
        _assert_bit_exact(WL, refl.shape, refl.indptr, refl.indices, refl.data, what='linear trial %d' % trial)


def test_keycompile_long_rows_use_bitmap_and_global_paths():
    """Rows longer than the shared-memory sort budget (8192): dense fc rows (bitmap ranking) and a very wide
    matrix (in-place global bitonic)."""
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(9)
    for (n_out, n_in) in [(5, 20000), (3, 300000)]:
        Wl = rs.randn(n_out, n_in).astype(np.float32); Wl[rs.rand(n_out, n_in) < 0.05] = 0.0
        bl = rs.randn(n_out).astype(np.float32)
        Ainv = sparse.MonomialKey(np.concatenate([rs.permutation(n_in), [n_in]]))
        WL = sparse.keyed_linear(torch.from_numpy(Wl), torch.from_numpy(bl), None, Ainv)
        ref = ko.sort_indices(ko.key_compile(None, ko.linear_matrix(Wl, bl), ko.monomial_key(Ainv.perm, Ainv.scale)))
        _assert_bit_exact(WL, ref.shape, ref.indptr, ref.indices, ref.data, what=str((n_out, n_in)))


# ---------------------------------------------------------------------------------------------
# Whole keyed networks against the reference goldens
def _load_lenet(z, pth_prefix='weights.'):
    from keynet_b200 import nets
    net = nets.LeNet_AvgPool().eval()
    net.load_state_dict({k[len(pth_prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pth_prefix)})
    return net


@pytest.mark.parametrize('golden,kwargs', [
    ('lenet_cfg1.npz', dict(global_geometric='permutation')),
    ('lenet_cfg3.npz', dict(global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)),
])
def test_lenet_keynet_matches_reference(golden, kwargs):
    """Same seed -> same keys -> bit-identical canonical CSR for every layer; forward within rtol 1e-4 of the
    reference's keyed forward; decrypted logits equal the plain net's argmax (north star)."""
    from keynet_b200 import system
    z = gu.load(golden)
    net = _load_lenet(z)
    np.random.seed(0)
    (sensor, knet) = system.Keynet((1, 28, 28), net, **kwargs)
    # keys
    (A, Ainv) = sensor.keypair()
    (perm, scale) = gu.monomial_from_coo(z, 'sensor.A')
    assert np.array_equal(A.perm, perm) and np.array_equal(A.scale.view(np.uint32), scale.view(np.uint32))
    # compiled matrices
    layers = gu.jstr(z, 'layers')
    assert [k for (k, _) in knet.keyedlayers()] == layers
    for (k, L) in knet.keyedlayers():
        _assert_bit_exact(L.W, *gu.csr_arrays(z, 'layer.%s.W' % k), what=k)
    assert knet.num_parameters() == int(z['num_parameters'])
    # encrypt + forward (batch of 4), host tensors in / out like the reference
    x = torch.from_numpy(z['x'])
    xc = sensor.fromtensor(x).encrypt().astensor()
    assert not xc.is_cuda
    assert _close(xc.numpy(), z['x_cipher'])
    y = knet.forward(xc).reshape(x.shape[0], -1).numpy()
    assert _close(y, z['logits_keyed'])
    assert np.allclose(y, z['logits_plain'], atol=1e-4)
    assert np.array_equal(y.argmax(1), z['logits_plain'].argmax(1))
    # per-layer activations
    h = xc.cuda()
    for (k, m) in knet._keynet.named_children():
        h = m(h)
        if ('layer.%s.y' % k) in z.files:
            ref = z['layer.%s.y' % k]
            ref = np.maximum(ref, 0) if getattr(m, '_fused_relu', False) else ref
            assert _close(h.cpu().numpy(), ref), k
    # N=1 through the public API returns outshape like the reference (system.py:133)
    y1 = knet.forward(sensor.fromtensor(x[0:1]).encrypt().astensor())
    assert tuple(y1.shape) == (10, 1, 1) and _close(y1.numpy(), z['logits_keyed_n1'])
    # decrypt round trip
    xd = sensor.fromtensor(x).encrypt().decrypt().astensor()
    assert tuple(xd.shape) == tuple(x.shape) and np.allclose(xd.numpy(), x.numpy(), atol=1e-5)


def test_identity_and_public_sensor():
    from keynet_b200 import system, nets
    torch.manual_seed(1)
    net = nets.LeNet_AvgPool().eval()
    (sensor, knet) = system.IdentityKeynet((1, 28, 28), net)
    x = torch.randn(1, 1, 28, 28)
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).numpy().flatten()
    assert np.allclose(y, net(x).detach().numpy().flatten(), atol=1e-5)        # reference test_keynet.py:26
    ps = system.PublicKeyedSensor((1, 28, 28))
    with pytest.raises(ValueError):
        ps.fromtensor(x).encrypt()
    assert tuple(ps.fromtensor(x).tensor().shape) == (1, 785)
    with pytest.raises(ValueError):                                             # broken homogeneous coordinate
        bad = sensor.fromtensor(x).encrypt().astensor().clone(); bad[:, -1] = 2.0
        knet.forward(bad)


def test_challenge_known_answer():
    from keynet_b200 import sparse
    from keynet_b200.torch import linear_to_affine
    z = gu.load('challenge_kat.npz')
    relu_after = dict(conv1=True, pool1=False, conv2=True, pool2=False, fc1=True, fc2=True, fc3=False)
    h = torch.from_numpy(z['x_linear']).cuda()
    for k in gu.jstr(z, 'layers'):
        W = sparse.SparseMatrix(gu.csr_arrays(z, 'layer.%s.W' % k))
        h = W.torchdot(h.t(), relu=relu_after[k]).t()
    enc = linear_to_affine(h).cpu().numpy().flatten()
    assert np.allclose(enc, z['y_printed'], atol=1e-4)          # demo/challenge.ipynb cell 5
    assert _close(h.cpu().numpy(), z['y_reference'])


def _numpy_weights(net, seed):
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for (name, p) in net.named_parameters():
            fan_in = int(np.prod(p.shape[1:])) if p.ndim > 1 else int(p.shape[0])
            bound = 1.0 / np.sqrt(max(1, fan_in))
            p.copy_(torch.from_numpy(rs.uniform(-bound, bound, size=tuple(p.shape)).astype(np.float32)))
    return net


def test_allconvnet_cfg2_digests_and_logits():
    """CIFAR AllConvNet with hierarchical block-permutation keys (BASELINE config 2): every compiled layer's
    canonical CSR has the reference's sha256 (261.6 M nnz, bit-exact), and the keyed forward matches."""
    from keynet_b200 import system, nets
    g = gu.load('acn_cfg2.json')
    net = _numpy_weights(nets.AllConvNet(batchnorm=False), seed=0).eval()
    np.random.seed(0)
    (sensor, knet) = system.Keynet((3, 32, 32), net, global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1))
    assert knet.num_parameters() == g['num_parameters']
    for (k, L) in knet.keyedlayers():
        ref = g['layers'][k]
        assert list(L.W.shape) == ref['shape'] and L.nnz() == ref['nnz'], k
        assert gu.digest(*L.W.csr_arrays()) == ref['sha256'], k
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(g['x_seed']))
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(2, -1).numpy()
    assert _close(y, np.array(g['logits_keyed']))
    assert np.allclose(y, np.array(g['logits_plain']), atol=1e-4)
    assert np.array_equal(y.argmax(1), np.array(g['logits_plain']).argmax(1))


def test_vgg16_reduced_channel_twins_digests():
    """BASELINE config 5 on reduced-channel twins of VGG16 layers at full spatial size, permutation keys on
    both sides: canonical CSR sha256 equals the reference's."""
    from keynet_b200 import system, layer
    from torch import nn
    g = gu.load('vggtwin_cfg5.json')
    for (name, t) in g.items():
        rs = np.random.RandomState(t['seed'])
        np.random.seed(t['seed'] + 1)
        if t['kind'] == 'conv':
            (C, U, V) = t['inshape']
            m = nn.Conv2d(C, t['M'], 3, padding=1)
            with torch.no_grad():
                m.weight.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.weight.shape)).astype(np.float32)))
                m.bias.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.bias.shape)).astype(np.float32)))
            (inshape, outshape) = (tuple(t['inshape']), (t['M'], U, V))
        elif t['kind'] == 'pool':
            (C, U, V) = t['inshape']
            m = nn.AvgPool2d(3, 2, 0, ceil_mode=True)
            (inshape, outshape) = (tuple(t['inshape']), (C, U // 2, V // 2))
        else:
            m = nn.Linear(t['infeat'], t['outfeat'])
            with torch.no_grad():
                m.weight.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.weight.shape)).astype(np.float32)))
                m.bias.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.bias.shape)).astype(np.float32)))
            (inshape, outshape) = ((t['infeat'], 1, 1), (t['outfeat'], 1, 1))
        (Ain, Ain_inv) = system.keygen(inshape, 'permutation', 'identity', 'identity', 'identity')
        (Aout, Aout_inv) = system.keygen(outshape, 'permutation', 'identity', 'identity', 'identity')
        L = layer.KeyedLayer(m, inshape, outshape, Aout, Ain_inv)
        assert list(L.W.shape) == t['shape'] and L.nnz() == t['nnz'], name
        assert gu.digest(*L.W.csr_arrays()) == t['sha256'], name


def test_keyed_relu_after_batchnorm_and_merge():
    """AllConvNet(batchnorm=True): conv->bn fusion and the explicit keyed ReLU (reference test_keynet.py:243,
    identity keys) -- keyed forward equals the plain net."""
    from keynet_b200 import system, nets
    import warnings
    torch.manual_seed(3)
    net = nets.AllConvNet(batchnorm=True).eval()
    with torch.no_grad():
        for m in (net.conv3_bn, net.conv6_bn):
            m.running_mean.uniform_(-0.1, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.1, 0.1)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        (sensor, knet) = system.Keynet((3, 32, 32), net, global_geometric='permutation')
    x = torch.randn(3, 3, 32, 32)
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(3, -1).numpy()
    yp = net(x).detach().numpy()
    assert np.allclose(y, yp, atol=1e-4), np.abs(y - yp).max()
    assert np.array_equal(y.argmax(1), yp.argmax(1))


# ---------------------------------------------------------------------------------------------
# Pattern-grouped execution format (csrc/pgroup.cu)
@pytest.mark.parametrize('case', ['conv_perm_gain', 'conv_identity_s2', 'linear', 'pool'])
@pytest.mark.parametrize('N', [32, 100, 128, 516])
def test_pattern_groups_match_csr_and_oracle(case, N):
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(17)

    def key(n, permute, gain):
        perm = np.concatenate([rs.permutation(n - 1), [n - 1]]) if permute else np.arange(n)
        scale = np.concatenate([rs.rand(n - 1) + 0.5, [1.0]]).astype(np.float32) if gain else np.ones(n, dtype=np.float32)
        return sparse.MonomialKey(perm, scale)
    if case == 'conv_perm_gain':
        (C, U, V, M) = (5, 10, 12, 24)
        W = sparse.keyed_toeplitz_conv2d((C, U, V), rs.randn(M, C, 3, 3).astype(np.float32), rs.randn(M).astype(np.float32), 1,
                                         key(M * U * V + 1, True, True), key(C * U * V + 1, True, True).transpose())
        expect_G = {M}
    elif case == 'conv_identity_s2':
        (C, U, V, M) = (3, 16, 16, 100)
        W = sparse.keyed_toeplitz_conv2d((C, U, V), rs.randn(M, C, 3, 3).astype(np.float32), rs.randn(M).astype(np.float32), 2,
                                         None, sparse.sparse_identity_matrix(C * U * V + 1))
        expect_G = {M}
    elif case == 'linear':
        W = sparse.keyed_linear(torch.from_numpy(rs.randn(70, 333).astype(np.float32)), torch.from_numpy(rs.randn(70).astype(np.float32)),
                                key(71, True, False), key(334, True, True).transpose())
        expect_G = {70}
    else:
        W = sparse.keyed_toeplitz_avgpool2d((6, 12, 12), 3, 2, key(6 * 36 + 1, True, False), key(6 * 144 + 1, True, False).transpose())
        expect_G = set()
    (ip, ix, dt) = W.csr_arrays()
    A = ko.csr(W.shape, ip, ix, dt)
    X = rs.randn(W.shape[1], N).astype(np.float32)
    X[-1] = 1.0                                      # homogeneous coordinate
    Xd = torch.from_numpy(X).cuda()
    W._pg = None
    y_csr = sparse.spmm(W, Xd, relu=True).cpu().numpy()
    W._pg = sparse.PatternGroups.build(W, min_group=4)
    if not expect_G:
        assert W._pg is None                         # pool rows all have distinct patterns: stays CSR
        return
    assert W._pg is not None and {c['G'] for c in W._pg.classes} == expect_G
    s = W._pg.summary()
    assert s['grouped_rows'] + s['rest_rows'] == W.shape[0]
    y_pg = sparse.spmm(W, Xd, relu=True).cpu().numpy()
    ref = ko.spmm(A, X, relu=True, threads=4)
    assert _close(y_pg, ref), np.abs(y_pg - ref).max()
    assert _close(y_pg, y_csr)
    assert bool((y_pg[-1] == 1.0).all())             # homogeneous row (rest CSR) stays exactly one


def test_pattern_groups_absent_for_unstructured_matrix():
    from keynet_b200 import sparse
    rs = np.random.RandomState(5)
    A = _rand_csr(rs, 300, 400, 0.05)
    W = sparse.SparseMatrix((A.shape, A.indptr, A.indices, A.data))
    assert sparse.PatternGroups.build(W) is None


# ---------------------------------------------------------------------------------------------
# tcgen05 (3xTF32) path of the pattern groups (csrc/pgroup_tc.cu)
@pytest.mark.parametrize('M,C,N', [(32, 3, 128), (96, 4, 256), (192, 2, 516), (100, 3, 384), (24, 5, 128), (512, 1, 260), (96, 20, 4096)])
def test_pattern_groups_tensor_core_path(M, C, N):
    """Same matrix through the fp32 CSR kernel, the fp32 grouped kernel and the tcgen05 kernel: all within
    rtol 1e-4 of the oracle (3xTF32 keeps ~fp32 accuracy; plain TF32 would not)."""
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(M + N)
    (U, V) = (6, 8)
    perm_gain = lambda n: sparse.MonomialKey(np.concatenate([rs.permutation(n - 1), [n - 1]]), np.concatenate([rs.rand(n - 1) + 0.5, [1.0]]).astype(np.float32))
    sparse.tensor_cores_enabled(True)
    sparse.PatternGroups.TC_MIN_K = 32                # exercise the tensor-core kernel also on these short reductions
    W = sparse.keyed_toeplitz_conv2d((C, U, V), rs.randn(M, C, 3, 3).astype(np.float32), rs.randn(M).astype(np.float32), 1,
                                     perm_gain(M * U * V + 1), perm_gain(C * U * V + 1).transpose())
    W._pg = sparse.PatternGroups.build(W, min_group=4)
    assert W._pg is not None
    uses_tc = [c['tc'] is not None for c in W._pg.classes]
    assert all(uses_tc) == (M >= sparse.PatternGroups.TC_MIN_G)
    (ip, ix, dt) = W.csr_arrays()
    X = (rs.randn(W.shape[1], N) * rs.choice([1e-3, 1.0, 50.0], size=(W.shape[1], 1))).astype(np.float32)
    X[-1] = 1.0
    Xd = torch.from_numpy(X).cuda()
    ref = ko.spmm(ko.csr(W.shape, ip, ix, dt), X, relu=True, threads=8)
    y_tc = sparse.spmm(W, Xd, relu=True).cpu().numpy()
    sparse.tensor_cores_enabled(False)
    try:
        y_simt = sparse.spmm(W, Xd, relu=True).cpu().numpy()
    finally:
        sparse.tensor_cores_enabled(True)
    assert _close(y_simt, ref), np.abs(y_simt - ref).max()
    assert _close(y_tc, ref), (np.abs(y_tc - ref).max(), np.abs(ref).max())
    # no ReLU variant
    y2 = sparse.spmm(W, Xd, relu=False).cpu().numpy()
    sparse.PatternGroups.TC_MIN_K = 128
    assert _close(y2, ko.spmm(ko.csr(W.shape, ip, ix, dt), X, relu=False, threads=8))


def test_unique_value_blocks_under_permutation_keys():
    """Permutation keys only relabel rows/columns, so all interior output pixels share one value block (the reference's
    unique tiles, but also under a GLOBAL permutation where its TiledMatrix cannot be used, system.py:360): a 3x3 conv
    has 9 distinct blocks (interior, 4 edges, 4 corners); gain keys make every block different."""
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(4)
    (C, U, V, M, N) = (4, 9, 11, 32, 256)
    f = rs.randn(M, C, 3, 3).astype(np.float32); b = rs.randn(M).astype(np.float32)
    perm = lambda n: sparse.MonomialKey(np.concatenate([rs.permutation(n - 1), [n - 1]]))
    gain = lambda n: sparse.MonomialKey(np.arange(n), np.concatenate([rs.rand(n - 1) + 0.5, [1.0]]).astype(np.float32))
    X = rs.randn(C * U * V + 1, N).astype(np.float32); X[-1] = 1.0
    for (A, Ainv, expect_unique) in [(None, perm(C * U * V + 1), 9), (perm(M * U * V + 1), perm(C * U * V + 1), None), (None, gain(C * U * V + 1), U * V)]:
        W = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, 1, A, Ainv)
        assert W._pg is not None and len(W._pg.classes) == 1
        s = W._pg.summary()
        assert s['classes'][0][0] == M and s['classes'][0][2] == U * V
        if expect_unique is not None:
            assert s['unique_blocks'][0] == expect_unique, s
        (ip, ix, dt) = W.csr_arrays()
        ref = ko.spmm(ko.csr(W.shape, ip, ix, dt), X, relu=True, threads=8)
        for tc in (True, False):
            sparse.tensor_cores_enabled(tc)
            try:
                y = sparse.spmm(W, torch.from_numpy(X).cuda(), relu=True).cpu().numpy()
            finally:
                sparse.tensor_cores_enabled(True)
            assert _close(y, ref), (tc, np.abs(y - ref).max())


# ---------------------------------------------------------------------------------------------
# TiledMatrix / Conv2dTiledMatrix structure (reference keynet/sparse.py:517-835, test/test_sparse.py:122)
def test_tiled_matrix_structure_matches_reference():
    from keynet_b200 import sparse, tiled
    z = gu.load('tiled_kat.npz')
    # avgpool (2,8,8), identity keys, tiles (4,8)
    W = sparse.keyed_toeplitz_avgpool2d((2, 8, 8), 3, 2, sparse.sparse_identity_matrix(2 * 16 + 1), sparse.sparse_identity_matrix(2 * 64 + 1))
    _assert_bit_exact(W, *gu.csr_arrays(z, 'pool.W'), what='pool W')
    T = tiled.TiledMatrix(W, tuple(int(t) for t in z['pool.tileshape']))
    assert np.array_equal(np.array(T.blocks(), dtype=np.int64), z['pool.blocks'])
    assert len(T.tiles()) == int(z['pool.ntiles']) and T.nnz() == int(z['pool.nnz'])
    E = T.tocsr(); E.sort_indices()
    (shape, ip, ix, dt) = gu.csr_arrays(z, 'pool.expanded')
    assert E.shape == shape and np.array_equal(E.indptr, ip) and np.array_equal(E.indices, ix) and np.array_equal(E.data, dt)
    assert sum(t.nnz for t in T.tiles()) == T.nnz() and T.tileshape() == (4, 8)
    # conv (3,8,8)->(4,8,8), identity keys, tiles (4,4), with bias column
    inshape = tuple(int(s) for s in z['conv.inshape']); outshape = tuple(int(s) for s in z['conv.outshape'])
    Wc = sparse.keyed_toeplitz_conv2d(inshape, z['conv.f'], z['conv.b'], 1, sparse.sparse_identity_matrix(int(np.prod(outshape)) + 1), sparse.sparse_identity_matrix(int(np.prod(inshape)) + 1))
    _assert_bit_exact(Wc, *gu.csr_arrays(z, 'conv.W'), what='conv W')
    Tc = tiled.Conv2dTiledMatrix(Wc, inshape, outshape, tuple(int(t) for t in z['conv.tileshape']), bias=True, sanitycheck=False)
    assert np.array_equal(np.array(Tc.blocks(), dtype=np.int64), z['conv.blocks'])
    assert Tc.nnz() == int(z['conv.nnz']) and Tc._n_tile_entries == int(z['conv.ntile_entries'])
    y = Tc.torchdot(torch.from_numpy(z['conv.x'])).numpy()
    assert _close(y, z['conv.y'])
    with pytest.raises(AssertionError):
        Tc.torchdot(torch.zeros(5, 2))


def test_tiled_keynets_match_plain_net():
    """reference test_keynet.py:37 (tiled identity keynet) and TiledPermutationKeynet: keyed == plain, atol 1e-5."""
    from keynet_b200 import system, nets, tiled
    torch.manual_seed(2)
    net = nets.LeNet_AvgPool().eval()
    x = torch.randn(3, 1, 28, 28)
    yp = net(x).detach().numpy()
    (sensor, knet) = system.Keynet((1, 28, 28), net, tileshape=(28, 28))
    assert any(isinstance(L.W, tiled.Conv2dTiledMatrix) for (k, L) in knet.keyedlayers())
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(3, -1).numpy()
    assert np.allclose(y, yp, atol=1e-5)
    (s0, k0) = system.Keynet((1, 28, 28), net)
    assert knet.num_parameters() < k0.num_parameters()          # unique-tile storage is smaller than the expanded matrices
    np.random.seed(3)
    (sensor, knet) = system.TiledPermutationKeynet((1, 28, 28), net, 4)
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(3, -1).numpy()
    assert np.allclose(y, yp, atol=1e-5)
    # local (block-repeated) permutation keys keep the matrices tile compressible
    assert knet.num_parameters() < k0.num_parameters()


def test_save_and_load_compiled_keynet(tmp_path):
    """a compiled keynet round-trips through a flat tensor file (no pickled classes, no recompilation)."""
    from keynet_b200 import system, nets, io
    torch.manual_seed(5)
    net = nets.LeNet_AvgPool().eval()
    np.random.seed(5)
    (sensor, knet) = system.Keynet((1, 28, 28), net, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    x = torch.randn(40, 1, 28, 28)
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(40, -1).numpy()
    p = io.save(str(tmp_path / 'lenet.keynet'), sensor, knet)
    (s2, k2) = io.load(p)
    assert [k for (k, _) in k2.keyedlayers()] == [k for (k, _) in knet.keyedlayers()] and k2.num_parameters() == knet.num_parameters()
    for ((_, a), (_, b)) in zip(knet.keyedlayers(), k2.keyedlayers()):
        _assert_bit_exact(b.W, a.W.shape, *a.W.csr_arrays())
    y2 = k2.forward(s2.fromtensor(x).encrypt().astensor()).reshape(40, -1).numpy()
    # the loaded network rebuilds its pattern groups from the CSR (columns ascending) while the compiled one lists a group's
    # columns in Toeplitz tap order: same matrices, different fp32 summation order at batch >= 32 ...
    assert np.allclose(y, y2, rtol=1e-5, atol=1e-6 * np.abs(y).max())
    # ... and bit-identical results on the CSR kernel (small batch)
    y8 = knet.forward(sensor.fromtensor(x[:8]).encrypt().astensor()).reshape(8, -1).numpy()
    assert np.array_equal(y8, k2.forward(s2.fromtensor(x[:8]).encrypt().astensor()).reshape(8, -1).numpy())
    # public release: keys removed, the keyed network still runs on ciphertext
    xc = sensor.fromtensor(x).encrypt().astensor()
    (s3, k3) = io.load(io.save(str(tmp_path / 'public.keynet'), None, knet.public()))
    assert s3 is None and np.allclose(k3.forward(xc).reshape(40, -1).numpy(), y, rtol=1e-5, atol=1e-6 * np.abs(y).max())


# ---------------------------------------------------------------------------------------------
# photometric bias / affine keys (keys with a bias column)
def test_keycompile_bias_keys_vs_oracle():
    """A = [[D P, b],[0, 1]] on both sides: indices bit-exact, every value bit-exact except the last column, which is a
    row reduction (fp32 rounding; scipy sums it in its own traversal order)."""
    from keynet_b200 import sparse
    ko = _ko()
    rs = np.random.RandomState(21)
    (C, U, V, M) = (3, 6, 8, 5)
    f = rs.randn(M, C, 3, 3).astype(np.float32); b = rs.randn(M).astype(np.float32)
    (R, K) = (M * U * V + 1, C * U * V + 1)

    def affine_key(n, permute):
        perm = np.concatenate([rs.permutation(n - 1), [n - 1]]) if permute else np.arange(n)
        scale = np.concatenate([rs.rand(n - 1) + 0.5, [1.0]]).astype(np.float32)
        bias = np.concatenate([rs.rand(n - 1), [0.0]]).astype(np.float32)
        return sparse.MonomialKey(perm, scale, bias)
    for (A, Ainv) in [(affine_key(R, True), affine_key(K, True)), (None, affine_key(K, False)), (affine_key(R, False), sparse.MonomialKey(np.arange(K)))]:
        W = sparse.keyed_toeplitz_conv2d((C, U, V), f, b, 1, A, Ainv)
        ref = ko.sort_indices(ko.key_compile(None if A is None else ko.csr_from_dense(A.todense()), ko.toeplitz_conv2d((C, U, V), f, b, 1), ko.csr_from_dense(Ainv.todense())))
        (ip, ix, dt) = W.csr_arrays()
        assert np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices)
        last = ix == K - 1
        assert np.array_equal(dt[~last].view(np.uint32), ref.data[~last].view(np.uint32))
        assert np.allclose(dt[last], ref.data[last], rtol=1e-5, atol=1e-6)
        assert last.sum() >= R - 1                        # the bias column is dense


def test_photometric_keynets_match_plain_net():
    """reference test/test_keynet.py:65-81: gain / bias / affine global photometric keys, keyed == plain."""
    from keynet_b200 import system, nets
    torch.manual_seed(7)
    net = nets.LeNet_AvgPool().eval()
    x = torch.randn(4, 1, 28, 28)
    yp = net(x).detach().numpy()
    for (kw, atol) in [(dict(global_photometric='uniform_random_gain', beta=1.0), 1e-5),
                       (dict(global_photometric='uniform_random_bias', gamma=1.0), 1e-5),
                       (dict(global_photometric='uniform_random_affine', beta=1.0, gamma=1.0), 1e-4),
                       (dict(global_geometric='permutation', global_photometric='uniform_random_affine', beta=1.0, gamma=1.0), 1e-4),
                       (dict(global_photometric='linear_bias', gamma=2.0), 1e-4)]:
        np.random.seed(1)
        (sensor, knet) = system.Keynet((1, 28, 28), net, **kw)
        xc = sensor.fromtensor(x).encrypt().astensor()
        y = knet.forward(xc).reshape(4, -1).numpy()
        assert np.allclose(y, yp, atol=atol), (kw, np.abs(y - yp).max())
        xd = sensor.decrypt().astensor()
        assert np.allclose(xd.numpy(), x.numpy(), atol=1e-4)
        if 'bias' in kw['global_photometric'] or 'affine' in kw['global_photometric']:
            assert not np.allclose(xc.numpy()[:, :-1], x.reshape(4, -1).numpy(), atol=1e-3)      # the image really is keyed


def test_splitk_dense_layer_matches_oracle():
    """A dense fully connected layer is ONE pattern group; with a long reduction it runs as K slices + a reduce pass
    (kn_splitk_reduce_f32).  Same result as the oracle, with and without the fused ReLU, at a ragged K."""
    from keynet_b200 import sparse
    from keynet_b200.sparse import MonomialKey
    ko = _ko()
    rs = np.random.RandomState(0)
    (n_out, n_in, N) = (300, 5001, 256)
    Wt = torch.from_numpy(rs.randn(n_out, n_in).astype(np.float32) / 70.0)
    b = torch.from_numpy(rs.randn(n_out).astype(np.float32))
    A = MonomialKey(np.concatenate([rs.permutation(n_out), [n_out]]), np.concatenate([rs.rand(n_out) + 0.5, [1.0]]).astype(np.float32))
    Ainv = MonomialKey(np.concatenate([rs.permutation(n_in), [n_in]]))
    W = sparse.keyed_linear(Wt, b, A, Ainv)
    W.optimize()
    cls = [c for c in W._pg.classes if c.get('splitk') is not None]
    assert len(cls) == 1 and cls[0]['splitk']['S'] >= 2
    X = rs.randn(n_in + 1, N).astype(np.float32); X[-1] = 1
    (ip, ix, dt) = W.csr_arrays()
    for relu in (False, True):
        ref = ko.spmm(ko.csr(W.shape, ip, ix, dt), X, relu=relu)
        y = sparse.spmm(W, torch.from_numpy(X).cuda(), relu=relu).cpu().numpy()
        assert _close(y, ref), np.abs(y - ref).max()
        assert np.array_equal(y[-1], np.ones(N, dtype=np.float32))
