"""Serialisation of compiled keynets.  The reference pickles whole Python objects including scipy
matrices (`vipy.util.save((sensor, knet), 'x.pkl')`, test/test_keynet.py:106, demo/challenge.ipynb); here a keyed
network is a flat dict of tensors -- canonical CSR per layer (int64 indptr, int32 indices, fp32 data), layer names and
flags, sensor keys as (perm, scale[, bias]) or CSR for general keys -- written with torch.save, so it loads without this package's classes being
picklable and without recompiling any key."""
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import layer as _layer
from . import system as _system
from .sparse import SparseMatrix, MonomialKey, SparseKey

FORMAT_VERSION = 1


def _key_state(K):
    if K is None:
        return None
    if isinstance(K, SparseKey):          # general key: CSR
        return {'indptr': torch.from_numpy(K.indptr), 'indices': torch.from_numpy(K.indices), 'data': torch.from_numpy(K.data.astype(np.float32)), 'shape': tuple(K.shape)}
    st = {'perm': torch.from_numpy(K.perm), 'scale': torch.from_numpy(K.scale)}
    if K.bias is not None:
        st['bias'] = torch.from_numpy(K.bias)
    return st


def _key_load(s):
    if s is None:
        return None
    if 'indptr' in s:
        return SparseKey(s['indptr'].numpy(), s['indices'].numpy(), s['data'].numpy(), tuple(s['shape']))
    return MonomialKey(s['perm'].numpy(), s['scale'].numpy(), s['bias'].numpy() if 'bias' in s else None)


def state_dict(sensor, knet):
    layers = []
    for (name, m) in knet._keynet.named_children():
        if isinstance(m, _layer.KeyedLayer):
            W = m.W
            if W._data is None:
                raise ValueError('layer "%s" dropped its canonical CSR (keep_csr=False): it cannot be exported' % name)
            off = int(W._indptr[0].item())
            layers.append({'name': name, 'kind': 'keyed', 'shape': tuple(W.shape), 'indptr': (W._indptr - off).cpu(),
                           'indices': W._indices[off:off + W.nnz()].cpu() if off else W._indices.cpu(), 'data': W._data[off:off + W.nnz()].cpu() if off else W._data.cpu(),
                           'layertype': m._layertype, 'repr': m._repr, 'fused_relu': bool(m._fused_relu),
                           'inshape': tuple(m._inshape), 'outshape': tuple(m._outshape)})
        else:
            layers.append({'name': name, 'kind': 'relu'})
    (A, Ainv) = (None, None) if sensor is None else sensor.keypair()
    return {'format': 'keynet_b200', 'version': FORMAT_VERSION, 'inshape': None if sensor is None else tuple(sensor._inshape[1:]),
            'sensor': {'A': _key_state(A), 'Ainv': _key_state(Ainv)}, 'outshape': tuple(knet._outshape), 'layers': layers,
            'embeddingkey': _key_state(knet._embeddingkey), 'imagekey': _key_state(knet._imagekey)}


def save(path, sensor, knet):
    torch.save(state_dict(sensor, knet), path)
    return path


def load(path, optimize=True):
    """-> (sensor, knet) on the current CUDA device; the sensor is None for a public keynet saved without keys."""
    s = torch.load(path, map_location='cpu', weights_only=True)
    assert s.get('format') == 'keynet_b200' and s.get('version') == FORMAT_VERSION, 'not a keynet_b200 file'
    keyed = OrderedDict()
    for L in s['layers']:
        if L['kind'] == 'relu':
            keyed[L['name']] = _layer.FusedReLU()
            continue
        m = _layer.KeyedLayer.__new__(_layer.KeyedLayer)
        nn.Module.__init__(m)
        (m._layertype, m._repr, m._fused_relu) = (L['layertype'], L['repr'], L['fused_relu'])
        (m._inshape, m._outshape, m._tileshape, m._rows) = (L['inshape'], L['outshape'], None, None)
        m.W = SparseMatrix((L['shape'], L['indptr'], L['indices'], L['data']))
        if optimize:
            m.W.optimize()
        keyed[L['name']] = m
    knet = _system.KeyedModel.__new__(_system.KeyedModel)
    knet._keynet = nn.Sequential(keyed)
    knet._embeddingkey = _key_load(s['embeddingkey'])
    knet._imagekey = _key_load(s['imagekey'])
    knet._layernames = set(L['name'] for L in s['layers'])
    knet._outshape = tuple(s['outshape'])
    knet._netshape = None
    sensor = None
    if s['sensor']['A'] is not None:
        sensor = _system.KeyedSensor(tuple(s['inshape']), (_key_load(s['sensor']['A']), _key_load(s['sensor']['Ainv'])))
    return (sensor, knet)
