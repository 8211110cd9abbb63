#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference is pure Python; its one missing dependency (`vipy`) is replaced by the
small stub in tests/golden/_vipy_stub.  Everything written here is data (matrices, vectors,
digests) -- no reference source is copied.

    python tests/golden/make_golden.py [name ...]      # default: all

Goldens (all little-endian .npz / .json):
  toeplitz_kat.npz        reference sparse_toeplitz_conv2d / avgpool2d on the shapes of
                          test/test_sparse.py:223,251 (+ a multi-channel case)
  keygen_kat.npz          reference keygen() outputs (A, Ainv) for several option sets, incl. the general key families
                          (Givens-orthogonal, doubly stochastic, the TiledOrthogonalKeynet option set)
  blockpermute_kat.npz    reference hierarchical_block_permutation_matrix()
  lenet_cfg1.npz          np.random.seed(0); PermutationKeynet(LeNet_AvgPool) (SURVEY §8d cfg 1)
  lenet_cfg3.npz          np.random.seed(0); Keynet(permutation + uniform_random_gain) (cfg 3)
  lenet_givens.npz        the reference's LeNet orthogonal configuration (test/test_keynet.py:180-197): Givens local keys +
                          affine photometric keys + hierarchical rotation, block memory order -- general-key compile
  challenge_kat.npz       demo/keynet_challenge_lenet_10AUG20.{pkl,png} known-answer test
  acn_cfg2.json           AllConvNet hierarchical-permutation keynet: per-layer nnz + sha256
  vggtwin_cfg5.json       reduced-channel twins of VGG16 layers, permutation keys: nnz + sha256
  tiled_kat.npz           reference TiledMatrix / Conv2dTiledMatrix structure on small layers
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path[:0] = [os.path.join(HERE, '_vipy_stub'), REF]

import numpy as np
import scipy.sparse
import torch
from torch import nn

import keynet.globals
keynet.globals.verbose(False)
import keynet.sparse
import keynet.system
import keynet.mnist
import keynet.cifar10
import keynet.blockpermute
import keynet.torch


# ---------------------------------------------------------------------------------------
def canon(A):
    """Canonical CSR (sorted column indices, duplicates summed) as (indptr i64, indices i32, data f32/f64)."""
    A = scipy.sparse.csr_matrix(A).copy()
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data


def digest(indptr, indices, data):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(indptr, dtype='<i8').tobytes())
    h.update(np.ascontiguousarray(indices, dtype='<i4').tobytes())
    h.update(np.ascontiguousarray(data, dtype='<f4').tobytes())
    return h.hexdigest()


def put_csr(d, prefix, A, keep_stored_order=False):
    (ip, ix, dt) = canon(A)
    d[prefix + '.shape'] = np.array(A.shape, dtype=np.int64)
    d[prefix + '.indptr'] = ip
    d[prefix + '.indices'] = ix
    d[prefix + '.data'] = dt
    if keep_stored_order:
        # reference storage order (unsorted indices): needed to pin the CPU SpMM oracle bit-for-bit
        B = scipy.sparse.csr_matrix(A)
        d[prefix + '.stored_indices'] = B.indices.astype(np.int32)
        d[prefix + '.stored_data'] = B.data


def put_coo(d, prefix, A):
    A = scipy.sparse.coo_matrix(A)
    d[prefix + '.shape'] = np.array(A.shape, dtype=np.int64)
    d[prefix + '.row'] = A.row.astype(np.int32)
    d[prefix + '.col'] = A.col.astype(np.int32)
    d[prefix + '.data'] = A.data


def numpy_weights(net, seed):
    """Deterministic (numpy legacy RNG) kaiming-uniform-like init so weights are reproducible anywhere."""
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for (name, p) in net.named_parameters():
            fan_in = int(np.prod(p.shape[1:])) if p.ndim > 1 else int(p.shape[0])
            bound = 1.0 / np.sqrt(max(1, fan_in))
            p.copy_(torch.from_numpy(rs.uniform(-bound, bound, size=tuple(p.shape)).astype(np.float32)))
    return net


class KeygenRecorder(object):
    """Wrap keynet.system.keygen to record every (shape, kwargs, A, Ainv) in call order."""

    def __init__(self):
        self.calls = []
        self._orig = keynet.system.keygen

    def __enter__(self):
        def wrapped(shape, *a, **k):
            (A, Ainv) = self._orig(shape, *a, **k)
            self.calls.append((tuple(int(s) for s in shape), A, Ainv))
            return (A, Ainv)
        keynet.system.keygen = wrapped
        return self

    def __exit__(self, *a):
        keynet.system.keygen = self._orig
        return False


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **d)
    print('[make_golden] wrote %s (%.1f KB)' % (name, os.path.getsize(path) / 1024.0))


# ---------------------------------------------------------------------------------------
def g_toeplitz():
    d = {}
    rs = np.random.RandomState(7)
    cases = {
        'conv_s2': dict(inshape=(1, 8, 16), M=4, k=3, stride=2),    # test_sparse.py:223 shape
        'conv_s1': dict(inshape=(3, 6, 5), M=2, k=3, stride=1),
        'conv_k1': dict(inshape=(4, 5, 5), M=3, k=1, stride=1),
        'conv_k5': dict(inshape=(2, 9, 7), M=2, k=5, stride=1),
        'conv_tiny': dict(inshape=(2, 1, 1), M=2, k=3, stride=1),   # only the centre tap is ever valid
    }
    for (name, c) in cases.items():
        (C, U, V) = c['inshape']
        f = rs.randn(c['M'], C, c['k'], c['k']).astype(np.float32)
        f[0, 0, 0, 0] = 0.0          # explicit zero coefficient must survive (offset trick, sparse.py:184)
        f[-1, -1, -1, -1] = 1e-9     # tiny weight: collapses to exactly 0 after (w+off)-off
        b = rs.randn(c['M']).astype(np.float32)
        b[0] = 0.0
        A = keynet.sparse.sparse_toeplitz_conv2d(c['inshape'], f.copy(), bias=b.copy(), stride=c['stride'])
        d[name + '.inshape'] = np.array(c['inshape']); d[name + '.stride'] = np.array(c['stride'])
        d[name + '.f'] = f; d[name + '.b'] = b
        put_csr(d, name + '.W', A)
        # functional check data (test_sparse.py:223): W.x == conv2d(x)
        x = rs.randn(2, C, U, V).astype(np.float32)
        y = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(f), torch.from_numpy(b), stride=c['stride'], padding=c['k'] // 2)
        d[name + '.x'] = x; d[name + '.y'] = y.numpy()
    for (name, inshape, k, s) in [('pool_s2', (1, 8, 16), 3, 2), ('pool_c3', (3, 6, 6), 3, 2), ('pool_s1', (2, 5, 4), 3, 1)]:
        A = keynet.sparse.sparse_toeplitz_avgpool2d(inshape, (inshape[0], inshape[0], k, k), s)
        d[name + '.inshape'] = np.array(inshape); d[name + '.k'] = np.array(k); d[name + '.stride'] = np.array(s)
        put_csr(d, name + '.W', A)
    save('toeplitz_kat.npz', d)


def g_keygen():
    d = {}
    cases = {
        'perm': dict(shape=(2, 8, 8), global_geometric='permutation', local_geometric='identity', global_photometric='identity', local_photometric='identity'),
        'gain': dict(shape=(2, 8, 8), global_geometric='identity', local_geometric='identity', global_photometric='uniform_random_gain', local_photometric='identity', beta=1.0),
        'perm_gain': dict(shape=(1, 28, 28), global_geometric='permutation', local_geometric='identity', global_photometric='uniform_random_gain', local_photometric='identity', beta=1.0),
        'hier': dict(shape=(3, 32, 32), global_geometric='hierarchical_permutation', local_geometric='identity', global_photometric='identity', local_photometric='identity',
                     hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1)),
        'hier_small': dict(shape=(10, 8, 8), global_geometric='hierarchical_permutation', local_geometric='identity', global_photometric='identity', local_photometric='identity',
                           hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1)),
        'hrot': dict(shape=(1, 16, 16), global_geometric='hierarchical_rotation', local_geometric='identity', global_photometric='identity', local_photometric='identity',
                     hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0,)),
        'blockorder_perm': dict(shape=(1, 28, 28), global_geometric='permutation', local_geometric='identity', global_photometric='identity', local_photometric='identity',
                                memoryorder='block', blocksize=14),
        'local_perm': dict(shape=(2, 8, 8), global_geometric='identity', local_geometric='permutation', global_photometric='identity', local_photometric='identity', blocksize=4),
        'local_gain': dict(shape=(2, 8, 8), global_geometric='identity', local_geometric='identity', global_photometric='identity', local_photometric='uniform_random_gain', blocksize=4, beta=2.0),
        'fc_perm': dict(shape=(120, 1, 1), global_geometric='permutation', local_geometric='identity', global_photometric='identity', local_photometric='identity'),
        'bias': dict(shape=(2, 4, 4), global_geometric='identity', local_geometric='identity', global_photometric='uniform_random_bias', local_photometric='identity', gamma=1.0),
        'affine': dict(shape=(2, 4, 4), global_geometric='permutation', local_geometric='identity', global_photometric='uniform_random_affine', local_photometric='identity', beta=1.0, gamma=1.0),
        # general (several entries per row) keys
        'givens_local': dict(shape=(2, 8, 8), global_geometric='identity', local_geometric='givens_orthogonal', global_photometric='identity', local_photometric='identity', alpha=5.0, blocksize=4),
        'givens_global': dict(shape=(3, 8, 8), global_geometric='givens_orthogonal', local_geometric='identity', global_photometric='identity', local_photometric='identity', alpha=30),
        'doubly_stochastic': dict(shape=(2, 8, 8), global_geometric='identity', local_geometric='doubly_stochastic', global_photometric='identity', local_photometric='identity', alpha=3.0, blocksize=4),
        'orthogonal_tiled': dict(shape=(2, 16, 16), global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1), global_photometric='identity',
                                 local_geometric='givens_orthogonal', alpha=4, blocksize=4, local_photometric='uniform_random_affine', beta=0.1, gamma=100.0, memoryorder='block', tileshape=(4, 4)),
        'givens_fc': dict(shape=(12, 1, 1), global_geometric='identity', local_geometric='givens_orthogonal', global_photometric='identity', local_photometric='uniform_random_gain', alpha=3.0, beta=1.0, blocksize=4),
    }
    names = []
    for (name, kw) in cases.items():
        kw = dict(kw)
        shape = kw.pop('shape')
        np.random.seed(11)
        (A, Ainv) = keynet.system.keygen(shape, **kw)
        put_coo(d, name + '.A', A)
        put_coo(d, name + '.Ainv', Ainv)
        d[name + '.args'] = np.array(json.dumps(dict(shape=shape, **kw)))
        names.append(name)
    d['names'] = np.array(json.dumps(names))
    save('keygen_kat.npz', d)


def g_blockpermute():
    d = {}
    cases = {
        'p01_32x32x3': dict(imgshape=(32, 32, 3), blockshape=(2, 2), permute_at_level=[0, 1], twist=False),
        'p0_16x16x1': dict(imgshape=(16, 16, 1), blockshape=(2, 2), permute_at_level=[0], twist=False),
        'p012_64x64x2': dict(imgshape=(64, 64, 2), blockshape=(2, 2), permute_at_level=[0, 1, 2], twist=False),
        't01_32x32x1': dict(imgshape=(32, 32, 1), blockshape=(2, 2), permute_at_level=[0, 1], twist=True),
        'p1_32x32x1': dict(imgshape=(32, 32, 1), blockshape=(2, 2), permute_at_level=[1], twist=False),
        'p01_27x27x1_b3': dict(imgshape=(27, 27, 1), blockshape=(3, 3), permute_at_level=[0, 1], twist=False),
    }
    names = []
    for (name, kw) in cases.items():
        P = keynet.blockpermute.hierarchical_block_permutation_matrix(kw['imgshape'], kw['blockshape'], kw['permute_at_level'], min_blocksize=8, seed=42, twist=kw['twist'], strict=False)
        P = scipy.sparse.coo_matrix(P)
        order = np.argsort(P.row)
        d[name + '.cols'] = P.col[order].astype(np.int64)   # P[r, cols[r]] = 1
        d[name + '.args'] = np.array(json.dumps(kw))
        names.append(name)
    d['names'] = np.array(json.dumps(names))
    save('blockpermute_kat.npz', d)


def _lenet(pth):
    net = keynet.mnist.LeNet_AvgPool()
    net.load_state_dict(torch.load(os.path.join(REF, 'models', pth), map_location='cpu'))
    net.eval()
    return net


def _record_keynet(d, net, inshape, f_make, N=4, xseed=0):
    """Run reference keying + forward, store weights, keys, per-layer canonical CSR and activations."""
    for (k, v) in net.state_dict().items():
        d['weights.' + k] = v.detach().cpu().numpy()
    with KeygenRecorder() as rec:
        (sensor, knet) = f_make(net)
    d['n_keygen_calls'] = np.array(len(rec.calls))
    for (i, (shape, A, Ainv)) in enumerate(rec.calls):
        d['keygen.%d.shape' % i] = np.array(shape)
        put_coo(d, 'keygen.%d.A' % i, A)
        put_coo(d, 'keygen.%d.Ainv' % i, Ainv)
    (A, Ainv) = sensor.keypair()
    put_coo(d, 'sensor.A', A)
    put_coo(d, 'sensor.Ainv', Ainv)

    x = torch.randn(N, *inshape, generator=torch.Generator().manual_seed(xseed))
    d['x'] = x.numpy()
    xc = sensor.fromtensor(x).encrypt().astensor()
    d['x_cipher'] = xc.numpy()
    names = []
    y = xc
    for (k, c) in knet._keynet.named_children():
        y = c.forward(y)
        if isinstance(c, keynet.layer.KeyedLayer):
            names.append(k)
            put_csr(d, 'layer.%s.W' % k, c.W._matrix, keep_stored_order=True)
            d['layer.%s.y' % k] = y.detach().numpy().astype(np.float32)   # pre-ReLU output of the keyed layer [N, R]
    d['layers'] = np.array(json.dumps(names))
    d['y_cipher'] = y.detach().numpy().astype(np.float32)
    d['logits_keyed'] = keynet.torch.linear_to_affine(y).detach().numpy()
    d['logits_plain'] = net.forward(x).detach().numpy()
    # single-image path through the public API (KeyedModel.forward is batch-1 only, torch.py:77)
    y1 = knet.forward(sensor.fromtensor(x[0:1]).encrypt().astensor())
    d['logits_keyed_n1'] = y1.detach().numpy()
    assert np.allclose(d['logits_keyed'], d['logits_plain'], atol=1e-4)
    return (sensor, knet)


def g_lenet_cfg1():
    d = {}
    net = _lenet('mnist_lenet_avgpool.pth')

    def make(net):
        np.random.seed(0)
        return keynet.system.PermutationKeynet((1, 28, 28), net, do_output_encryption=False)
    (sensor, knet) = _record_keynet(d, net, (1, 28, 28), make)
    d['num_parameters'] = np.array(knet.num_parameters())
    save('lenet_cfg1.npz', d)


def g_lenet_cfg3():
    d = {}
    net = _lenet('mnist_lenet_avgpool_fiberbundle.pth')

    def make(net):
        np.random.seed(0)
        return keynet.system.Keynet((1, 28, 28), net, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    (sensor, knet) = _record_keynet(d, net, (1, 28, 28), make)
    d['num_parameters'] = np.array(knet.num_parameters())
    save('lenet_cfg3.npz', d)


def g_lenet_givens():
    """General (non-monomial) keys: the reference's own LeNet orthogonal configuration (test/test_keynet.py:180-197):
    Givens-rotation local keys + affine photometric keys + hierarchical rotation, block memory order."""
    d = {}
    net = numpy_weights(keynet.mnist.LeNet_AvgPool(), 11).eval()

    def make(net):
        np.random.seed(0)
        return keynet.system.Keynet((1, 28, 28), net, tileshape=None,
                                    global_geometric='hierarchical_rotation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0),
                                    global_photometric='uniform_random_bias',
                                    local_geometric='givens_orthogonal', alpha=2.0, blocksize=8,
                                    local_photometric='uniform_random_affine', beta=1.0, gamma=1.0,
                                    memoryorder='block')
    (sensor, knet) = _record_keynet(d, net, (1, 28, 28), make, N=2)
    d['num_parameters'] = np.array(knet.num_parameters())
    for k in [k for k in d if k.startswith('keygen.')]:          # the sensor key pins the RNG stream; the rest is redundant
        del d[k]
    save('lenet_givens.npz', d)


def g_challenge():
    import pickle
    import PIL.Image
    d = {}
    with open(os.path.join(REF, 'demo', 'keynet_challenge_lenet_10AUG20.pkl'), 'rb') as f:
        (sensor, knet) = pickle.load(f)
    img = np.array(PIL.Image.open(os.path.join(REF, 'demo', 'keynet_challenge_lenet_10AUG20.png')))
    red = img[:, :, 0] if img.ndim == 3 else img
    x = torch.as_tensor(red.astype(np.float32) / 255.0).reshape(1, 1, 28, 28)
    xl = keynet.torch.affine_to_linear(x)
    d['x_linear'] = xl.numpy()
    names = []
    y = xl
    for (k, c) in knet._keynet.named_children():
        y = c.forward(y)
        if isinstance(c, keynet.layer.KeyedLayer):
            names.append(k)
            M = c.W._matrix
            d['layer.%s.src_dtype' % k] = np.array(str(M.dtype))
            put_csr(d, 'layer.%s.W' % k, M.astype(np.float32))   # fp32 cast of the pickled (partly fp64) matrices
    d['layers'] = np.array(json.dumps(names))
    d['y_reference'] = y.detach().numpy().astype(np.float64)
    # demo/challenge.ipynb cell 5 printed output (4 d.p.)
    d['y_printed'] = np.array([-0.0592, -0.0604, 0.0438, -0.0802, 0.0204, 0.0233, -0.0330, 0.0081, 0.0433, -0.0841])
    enc = keynet.torch.linear_to_affine(y).detach().numpy().flatten()
    assert np.allclose(enc, d['y_printed'], atol=1e-4), enc
    save('challenge_kat.npz', d)


def _layer_digests(knet):
    out = {}
    for (k, c) in knet._keynet.named_children():
        if isinstance(c, keynet.layer.KeyedLayer):
            (ip, ix, dt) = canon(c.W._matrix)
            out[k] = dict(shape=[int(s) for s in c.W.shape], nnz=int(c.nnz()), sha256=digest(ip, ix, dt),
                          data_sum=float(np.sum(dt.astype(np.float64))), indices_sum=int(np.sum(ix.astype(np.int64))))
    return out


def g_acn_cfg2():
    t0 = time.time()
    net = numpy_weights(keynet.cifar10.AllConvNet(batchnorm=False), seed=0).eval()
    np.random.seed(0)
    with KeygenRecorder() as rec:
        (sensor, knet) = keynet.system.Keynet((3, 32, 32), net, global_geometric='hierarchical_permutation',
                                              hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1))
    out = dict(config='np.random.seed(0); Keynet((3,32,32), AllConvNet(batchnorm=False) w/ numpy_weights(seed=0), '
                      'global_geometric=hierarchical_permutation, hierarchical_blockshape=(2,2), hierarchical_permute_at_level=(0,1))',
               layers=_layer_digests(knet), num_parameters=int(knet.num_parameters()))
    (A, Ainv) = sensor.keypair()
    A = scipy.sparse.csr_matrix(A); A.sort_indices()
    out['sensor_perm_sha256'] = hashlib.sha256(A.indices.astype('<i4').tobytes()).hexdigest()
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    y = knet._keynet.forward(sensor.fromtensor(x).encrypt().astensor())
    out['x_seed'] = 0
    out['logits_keyed'] = keynet.torch.linear_to_affine(y).detach().numpy().astype(np.float64).tolist()
    out['logits_plain'] = net.forward(x).detach().numpy().astype(np.float64).tolist()
    with open(os.path.join(HERE, 'acn_cfg2.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('[make_golden] wrote acn_cfg2.json in %.0f s' % (time.time() - t0))


def g_vggtwin():
    """Reduced-channel twins of VGG16 keyed layers at full spatial size, global permutation keys on both sides."""
    out = {}
    twins = {
        'conv1_1_twin': dict(kind='conv', inshape=(3, 224, 224), M=4),
        'conv1_2_twin': dict(kind='conv', inshape=(4, 224, 224), M=4),
        'pool1_2_twin': dict(kind='pool', inshape=(4, 224, 224)),
        'conv3_1_twin': dict(kind='conv', inshape=(16, 56, 56), M=16),
        'conv5_1_twin': dict(kind='conv', inshape=(64, 14, 14), M=64),
        'fc_twin': dict(kind='fc', infeat=64 * 7 * 7, outfeat=256),
    }
    for (name, t) in twins.items():
        rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(ord(c) for c in name))
        np.random.seed(sum(ord(c) for c in name) + 1)
        if t['kind'] == 'conv':
            (C, U, V) = t['inshape']
            m = nn.Conv2d(C, t['M'], 3, padding=1)
            with torch.no_grad():
                m.weight.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.weight.shape)).astype(np.float32)))
                m.bias.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.bias.shape)).astype(np.float32)))
            outshape = (t['M'], U, V); inshape = t['inshape']
        elif t['kind'] == 'pool':
            (C, U, V) = t['inshape']
            m = nn.AvgPool2d(3, 2, 0, ceil_mode=True)
            outshape = (C, U // 2, V // 2); inshape = t['inshape']
        else:
            m = nn.Linear(t['infeat'], t['outfeat'])
            with torch.no_grad():
                m.weight.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.weight.shape)).astype(np.float32)))
                m.bias.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(m.bias.shape)).astype(np.float32)))
            inshape = (t['infeat'], 1, 1); outshape = (t['outfeat'], 1, 1)
        (Ain, Ain_inv) = keynet.system.keygen(inshape, 'permutation', 'identity', 'identity', 'identity')
        (Aout, Aout_inv) = keynet.system.keygen(outshape, 'permutation', 'identity', 'identity', 'identity')
        L = keynet.layer.KeyedLayer(m, inshape, outshape, Aout, Ain_inv)
        (ip, ix, dt) = canon(L.W._matrix)
        out[name] = dict(t, seed=sum(ord(c) for c in name), shape=[int(s) for s in L.W.shape], nnz=int(L.nnz()), sha256=digest(ip, ix, dt))
        print('   ', name, out[name]['shape'], out[name]['nnz'])
    with open(os.path.join(HERE, 'vggtwin_cfg5.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('[make_golden] wrote vggtwin_cfg5.json')


def g_tiled():
    d = {}
    rs = np.random.RandomState(3)
    # TiledMatrix on a keyed avgpool (identity keys): blocks + unique tiles  (sparse.py:517)
    W = keynet.sparse.sparse_toeplitz_avgpool2d((2, 8, 8), (2, 2, 3, 3), 2)
    W = scipy.sparse.eye(W.shape[0], dtype=np.float32).tocsr().dot(W).dot(scipy.sparse.eye(W.shape[1], dtype=np.float32).tocsr())
    T = keynet.sparse.TiledMatrix(W, (4, 8))
    put_csr(d, 'pool.W', W)
    d['pool.tileshape'] = np.array((4, 8))
    d['pool.blocks'] = np.array(T._blocks, dtype=np.int64)
    d['pool.ntiles'] = np.array(len(T._tiles))
    d['pool.nnz'] = np.array(T.nnz())
    put_csr(d, 'pool.expanded', T.tocsr())
    # Conv2dTiledMatrix on a keyed conv (identity keys)  (sparse.py:690)
    inshape = (3, 8, 8); outshape = (4, 8, 8)
    f = rs.randn(4, 3, 3, 3).astype(np.float32); b = rs.randn(4).astype(np.float32)
    W = keynet.sparse.sparse_toeplitz_conv2d(inshape, f.copy(), bias=b.copy(), stride=1)
    W = scipy.sparse.eye(W.shape[0], dtype=np.float32).tocsr().dot(W).dot(scipy.sparse.eye(W.shape[1], dtype=np.float32).tocsr())
    T = keynet.sparse.Conv2dTiledMatrix(W, inshape, outshape, (4, 4), bias=True, sanitycheck=False)
    d['conv.f'] = f; d['conv.b'] = b
    d['conv.inshape'] = np.array(inshape); d['conv.outshape'] = np.array(outshape); d['conv.tileshape'] = np.array((4, 4))
    put_csr(d, 'conv.W', W)
    d['conv.blocks'] = np.array(T._blocks, dtype=np.int64)
    d['conv.nnz'] = np.array(T.nnz())
    d['conv.ntile_entries'] = np.array(len(T._tiles))
    put_csr(d, 'conv.expanded', T.tocsr())
    x = rs.randn(W.shape[1], 3).astype(np.float32)
    d['conv.x'] = x
    d['conv.y'] = T.torchdot(torch.from_numpy(x)).numpy()
    save('tiled_kat.npz', d)


def g_params():
    """Rows of the paper's parameter-count table (demo/figures.py:236-293) produced by the unmodified reference: numpy-seeded
    weights (seed 0), np.random.seed(0) before every factory call.  LeNet: every row; AllConvNet: identity / permutation /
    TiledPermutationKeynet-8 (the other rows take the reference minutes each)."""
    import keynet.mnist, keynet.cifar10, keynet.system
    keynet.globals.verbose(False)
    rows = []

    def keyed(f, *a, **k):
        np.random.seed(0)
        (sensor, knet) = f(*a, **k)
        return int(knet.num_parameters())
    net = numpy_weights(keynet.mnist.LeNet_AvgPool(), 0).eval()
    inshape = (1, 28, 28)
    rows.append(['lenet', int(keynet.torch.count_parameters(net))])
    rows.append(['IdentityKeynet (lenet)', keyed(keynet.system.IdentityKeynet, inshape, net)])
    rows.append(['PermutationKeynet (lenet)', keyed(keynet.system.PermutationKeynet, inshape, net)])
    for k in (2, 4, 8):
        rows.append(['TiledPermutationKeynet-%d (lenet)' % k, keyed(keynet.system.TiledPermutationKeynet, inshape, net, k)])
    net = numpy_weights(keynet.cifar10.AllConvNet(batchnorm=False), 0).eval()
    inshape = (3, 32, 32)
    rows.append(['allconvnet', int(keynet.torch.count_parameters(net))])
    rows.append(['IdentityKeynet (allconvnet)', keyed(keynet.system.IdentityKeynet, inshape, net)])
    rows.append(['PermutationKeynet (allconvnet)', keyed(keynet.system.PermutationKeynet, inshape, net)])
    rows.append(['TiledPermutationKeynet-8 (allconvnet)', keyed(keynet.system.TiledPermutationKeynet, inshape, net, 8)])
    with open(os.path.join(HERE, 'params_kat.json'), 'w') as f:
        json.dump(rows, f, indent=1)


ALL = dict(params=g_params, toeplitz=g_toeplitz, keygen=g_keygen, blockpermute=g_blockpermute, lenet_cfg1=g_lenet_cfg1, lenet_cfg3=g_lenet_cfg3,
           challenge=g_challenge, lenet_givens=g_lenet_givens, acn_cfg2=g_acn_cfg2, vggtwin=g_vggtwin, tiled=g_tiled)

if __name__ == '__main__':
    which = sys.argv[1:] or list(ALL.keys())
    for w in which:
        t0 = time.time()
        ALL[w]()
        print('[make_golden] %s done in %.1f s' % (w, time.time() - t0))
