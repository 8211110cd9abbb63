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


# =============================================================================================
# Reference pickles: `vipy.util.save((sensor, knet), 'x.pkl')` in the reference writes a (dill-flavoured)
# pickle of its own classes holding scipy matrices (test/test_keynet.py:106,144,166,191; demo/keynet_challenge_lenet_10AUG20.pkl).
# It is read here WITHOUT importing the reference, scipy or dill and without executing anything the file asks for: a
# restricted unpickler maps every class the format uses to an inert record (its attribute dict) and allows exactly one
# callable, numpy's array reconstructor.  The records are then turned into device-resident layers of this package.
class _Record(object):
    """Inert stand-in for a pickled reference / scipy / torch object: just its state."""
    _kind = None

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):      # (dict, slots) form
            state = dict(state[0] or {}, **state[1])
        self.__dict__.update(state)


def _record_class(kind):
    return type('_Record_' + kind.replace('.', '_'), (_Record,), {'_kind': kind})


_ALLOWED_RECORDS = {
    ('keynet.system', 'PublicKeyedSensor'), ('keynet.system', 'KeyedSensor'), ('keynet.system', 'KeyedModel'), ('keynet.layer', 'KeyedLayer'),
    ('keynet.sparse', 'SparseMatrix'), ('keynet.sparse', 'TiledMatrix'), ('keynet.sparse', 'Conv2dTiledMatrix'), ('keynet.sparse', 'DiagonalTiledMatrix'),
    ('scipy.sparse.csr', 'csr_matrix'), ('scipy.sparse._csr', 'csr_matrix'), ('scipy.sparse.coo', 'coo_matrix'), ('scipy.sparse._coo', 'coo_matrix'),
    ('scipy.sparse.dia', 'dia_matrix'), ('scipy.sparse._dia', 'dia_matrix'), ('scipy.sparse.csc', 'csc_matrix'), ('scipy.sparse._csc', 'csc_matrix'),
    ('torch.nn.modules.container', 'Sequential'), ('torch.nn.modules.activation', 'ReLU'),
}


class _ModuleToken(object):
    def __init__(self, name):
        self.name = name


def _dill_import_module(name, *a):
    if name not in ('numpy.core._multiarray_umath', 'numpy._core._multiarray_umath', 'numpy.core.multiarray', 'numpy._core.multiarray'):
        raise ValueError('reference pickle asks for module "%s": refused' % name)
    return _ModuleToken(name)


def _dill_get_attr(obj, name):
    if isinstance(obj, _ModuleToken) and name == '_reconstruct':
        try:
            from numpy._core.multiarray import _reconstruct
        except ImportError:                                   # numpy 1.x
            from numpy.core.multiarray import _reconstruct
        return _reconstruct
    raise ValueError('reference pickle asks for attribute "%s": refused' % name)


def _dill_load_type(name):
    table = {'set': set, 'dict': dict, 'list': list, 'tuple': tuple, 'NoneType': type(None)}
    if name not in table:
        raise ValueError('reference pickle asks for type "%s": refused' % name)
    return table[name]


def _restricted_unpickler(f):
    import pickle
    import collections
    cache = {}

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            if (module, name) in _ALLOWED_RECORDS:
                return cache.setdefault((module, name), _record_class(module + '.' + name))
            if (module, name) == ('dill._dill', '_import_module'):
                return _dill_import_module
            if (module, name) == ('dill._dill', '_get_attr'):
                return _dill_get_attr
            if (module, name) == ('dill._dill', '_load_type'):
                return _dill_load_type
            if (module, name) == ('collections', 'OrderedDict'):
                return collections.OrderedDict
            if (module, name) == ('numpy', 'ndarray'):
                return np.ndarray
            if (module, name) == ('numpy', 'dtype'):
                return np.dtype
            if module in ('numpy.core.multiarray', 'numpy._core.multiarray') and name in ('_reconstruct', 'scalar'):
                return _dill_get_attr(_ModuleToken(module), '_reconstruct') if name == '_reconstruct' else np.core.multiarray.scalar
            raise ValueError('reference pickle references %s.%s: refused (only the classes of the keyed path are mapped)' % (module, name))
    return U(f)


def _record_to_csr(M):
    """scipy matrix record -> (shape, indptr, indices, data) with ascending columns; data keeps its dtype (fp32 / fp64)."""
    kind = M._kind.rsplit('.', 1)[-1]
    shape = tuple(int(v) for v in M._shape)
    if kind == 'csr_matrix':
        (indptr, indices, data) = (np.asarray(M.indptr, dtype=np.int64), np.asarray(M.indices, dtype=np.int64), np.asarray(M.data))
        rows = np.repeat(np.arange(shape[0]), np.diff(indptr))
    elif kind == 'coo_matrix':
        (rows, indices, data) = (np.asarray(M.row, dtype=np.int64), np.asarray(M.col, dtype=np.int64), np.asarray(M.data))
    elif kind == 'dia_matrix':
        (rs, cs, vs) = ([], [], [])
        for (d, off) in zip(np.atleast_2d(M.data), np.asarray(M.offsets).reshape(-1)):
            j = np.arange(max(0, off), min(shape[1], shape[0] + off))          # data[k, j] sits at A[j - off, j]
            rs.append(j - off); cs.append(j); vs.append(np.asarray(d)[j])
        (rows, indices, data) = (np.concatenate(rs), np.concatenate(cs), np.concatenate(vs))
        keep = data != 0
        (rows, indices, data) = (rows[keep], indices[keep], data[keep])
    elif kind == 'csc_matrix':
        cols = np.repeat(np.arange(shape[1]), np.diff(np.asarray(M.indptr)))
        (rows, indices, data) = (np.asarray(M.indices, dtype=np.int64), cols, np.asarray(M.data))
    else:
        raise ValueError('unsupported matrix record %s' % M._kind)
    order = np.lexsort((indices, rows))
    indptr = np.zeros(shape[0] + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return (shape, np.cumsum(indptr), indices[order].astype(np.int32), data[order])


def load_reference_pickle(path, optimize=True):
    """(sensor, knet) from a pickle written by the REFERENCE (vipy.util.save((sensor, knet), path)), e.g.
    demo/keynet_challenge_lenet_10AUG20.pkl.  Nothing of the reference, scipy or dill is imported; unknown classes are
    refused.  Matrices the reference stored in float64 (conv / pool layers of the challenge keynet) are converted to the
    float32 the B200 path computes in.  The sensor is a PublicKeyedSensor when the file holds one (no keys), else a
    KeyedSensor with the stored key matrices (as general SparseKeys)."""
    with open(path, 'rb') as f:
        obj = _restricted_unpickler(f).load()
    assert isinstance(obj, (tuple, list)) and len(obj) == 2, 'expected a (sensor, knet) pair'
    (rs, rk) = obj
    assert rk._kind == 'keynet.system.KeyedModel'
    keyed = OrderedDict()
    for (name, r) in rk._keynet._modules.items():
        if r._kind == 'keynet.layer.KeyedLayer':
            (shape, indptr, indices, data) = _record_to_csr(r.W._matrix)
            m = _layer.KeyedLayer.__new__(_layer.KeyedLayer)
            nn.Module.__init__(m)
            (m._layertype, m._repr, m._fused_relu) = (str(r._layertype), str(getattr(r, '_repr', name)), False)
            (m._inshape, m._outshape, m._tileshape, m._rows) = (getattr(r, '_inshape', None), getattr(r, '_outshape', None), None, None)
            m.W = SparseMatrix((shape, indptr, indices, data.astype(np.float32)))
            keyed[name] = m
        elif r._kind.endswith('ReLU'):
            # the reference runs an un-keyed nn.ReLU after the keyed layer (system.py:92): fused into that layer's epilogue here
            prev = next(reversed(keyed)) if len(keyed) else None
            assert prev is not None and isinstance(keyed[prev], _layer.KeyedLayer), 'ReLU without a keyed predecessor'
            keyed[prev].fuse_relu(True)
            keyed[name] = _layer.FusedReLU()
        else:
            raise ValueError('unsupported module %s in the pickled keynet' % r._kind)
    if optimize:
        for m in keyed.values():
            if isinstance(m, _layer.KeyedLayer):
                m.W.optimize()
    knet = _system.KeyedModel.__new__(_system.KeyedModel)
    knet._keynet = nn.Sequential(keyed)

    def key_of(M):
        if M is None:
            return None
        (shape, indptr, indices, data) = _record_to_csr(M)
        return SparseKey(indptr, indices, data.astype(np.float32), shape)
    knet._embeddingkey = key_of(getattr(rk, '_embeddingkey', None))
    knet._imagekey = key_of(getattr(rk, '_imagekey', None))
    knet._layernames = set(getattr(rk, '_layernames', set(keyed.keys())))
    out_rows = [m for m in keyed.values() if isinstance(m, _layer.KeyedLayer)][-1].W.shape[0] - 1
    knet._outshape = (out_rows, 1, 1)
    knet._netshape = None
    inshape = tuple(int(v) for v in rs._inshape)[-3:]
    if rs._kind.endswith('PublicKeyedSensor'):
        sensor = _system.PublicKeyedSensor(inshape)
    else:
        sensor = _system.KeyedSensor(inshape, (key_of(rs._encryptkey), key_of(rs._decryptkey)))
    return (sensor, knet)


# =============================================================================================
# Row-sharded keynets (dist.ShardedKeyedModel): every rank saves / loads its own shard file.
def save_shard(path, model):
    """One file per rank: this rank's rows of every keyed layer (canonical CSR over the gathered column layout), the
    shard bookkeeping and the sensor keys.  Needs keep_csr=True.  load_shard() restores it without recompiling."""
    layers = []
    for (name, L) in zip([k for (k, _) in model._model.keyedlayers()], model.layers):
        W = L.W
        if W._data is None:
            raise ValueError('layer "%s" holds no CSR (keep_csr=False): it cannot be exported' % name)
        off = int(W._indptr[0].item())
        sh = L._shard
        layers.append({'name': name, 'shape': tuple(W.shape), 'indptr': (W._indptr - off).cpu(), 'indices': W._indices[off:off + W.nnz()].cpu(), 'data': W._data[off:off + W.nnz()].cpu(),
                       'layertype': L._layertype, 'repr': L._repr, 'fused_relu': bool(L._fused_relu), 'chunk': int(sh.chunk), 'n_phys': int(sh.n_phys), 'n_rows': int(sh.n_rows),
                       'my_rows': torch.from_numpy(np.ascontiguousarray(sh.my_rows)), 'position': torch.from_numpy(np.ascontiguousarray(sh.position))})
    (A, Ainv) = model.sensor.keypair()
    torch.save({'format': 'keynet_b200.shard', 'version': FORMAT_VERSION, 'rank': model.rank, 'world': model.world, 'fused': model.fused, 'selective': model.selective,
                'inshape': tuple(model.sensor._inshape[1:]), 'outshape': tuple(model._outshape), 'sensor': {'A': _key_state(A), 'Ainv': _key_state(Ainv)}, 'layers': layers}, path)
    return path


def load_shard(path, rank, world, group=None, optimize=True):
    """-> dist.ShardedKeyedModel holding the shard saved by save_shard() for this (rank, world)."""
    from . import dist as _dist
    s = torch.load(path, map_location='cpu', weights_only=True)
    assert s.get('format') == 'keynet_b200.shard' and s.get('version') == FORMAT_VERSION, 'not a keynet_b200 shard file'
    assert (int(s['rank']), int(s['world'])) == (int(rank), int(world)), 'shard file belongs to rank %d of %d' % (s['rank'], s['world'])
    m = _dist.ShardedKeyedModel.__new__(_dist.ShardedKeyedModel)
    (m.rank, m.world, m.group, m.fused, m.selective) = (int(rank), int(world), group, bool(s['fused']), bool(s['selective']))
    (m._symm, m._masks, m.time_layers, m._events, m.flag_sync) = ({}, None, False, [], True)
    m.sensor = _system.KeyedSensor(tuple(s['inshape']), (_key_load(s['sensor']['A']), _key_load(s['sensor']['Ainv'])))
    m._outshape = tuple(s['outshape'])
    layers = []
    for L in s['layers']:
        k = _layer.KeyedLayer.__new__(_layer.KeyedLayer)
        nn.Module.__init__(k)
        (k._layertype, k._repr, k._fused_relu, k._tileshape, k._rows) = (L['layertype'], L['repr'], L['fused_relu'], None, None)
        k.W = SparseMatrix((L['shape'], L['indptr'], L['indices'], L['data']))
        if optimize:
            k.W.optimize()
        sh = _dist.LayerShard.__new__(_dist.LayerShard)
        (sh.chunk, sh.n_phys, sh.n_rows, sh.my_rows, sh.position) = (L['chunk'], L['n_phys'], L['n_rows'], L['my_rows'].numpy(), L['position'].numpy())
        k._shard = sh
        layers.append(k)
    m.layers = layers
    m._model = None
    return m
