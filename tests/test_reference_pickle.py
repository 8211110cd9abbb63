"""Import of the reference's own pickles: keynet_b200.io.load_reference_pickle reads
demo/keynet_challenge_lenet_10AUG20.pkl (fixture copy under tests/golden/reference_demo/) with a RESTRICTED unpickler -- the
reference, scipy and dill are not imported, unknown globals are refused -- and the imported keynet reproduces the encoding
printed in demo/challenge.ipynb cell 5."""
import io as _io
import os
import pickle

import numpy as np
import pytest

from tests import golden_util as gu

HERE = os.path.dirname(os.path.abspath(__file__))
PKL = os.path.join(HERE, 'golden', 'reference_demo', 'keynet_challenge_lenet_10AUG20.pkl')
PNG = os.path.join(HERE, 'golden', 'reference_demo', 'keynet_challenge_lenet_10AUG20.png')


def test_restricted_unpickler_yields_the_golden_matrices():
    from keynet_b200 import io
    z = gu.load('challenge_kat.npz')
    with open(PKL, 'rb') as f:
        (rs, rk) = io._restricted_unpickler(f).load()
    assert rs._kind == 'keynet.system.PublicKeyedSensor' and rk._kind == 'keynet.system.KeyedModel'
    names = [k for (k, r) in rk._keynet._modules.items() if r._kind.endswith('KeyedLayer')]
    assert names == ['conv1', 'pool1', 'conv2', 'pool2', 'fc1', 'fc2', 'fc3']
    for k in names:
        (shape, ip, ix, dt) = io._record_to_csr(rk._keynet._modules[k].W._matrix)
        assert tuple(z['layer.%s.W.shape' % k]) == shape
        assert np.array_equal(z['layer.%s.W.indptr' % k], ip) and np.array_equal(z['layer.%s.W.indices' % k], ix)
        assert np.array_equal(z['layer.%s.W.data' % k].view(np.uint32), dt.astype(np.float32).view(np.uint32))
        assert str(dt.dtype) == str(z['layer.%s.src_dtype' % k])


def test_restricted_unpickler_refuses_anything_else():
    from keynet_b200 import io

    class Evil(object):
        def __reduce__(self):
            return (os.system, ('echo pwned',))
    import subprocess
    for payload in (pickle.dumps(Evil()), pickle.dumps(subprocess.Popen), pickle.dumps(_io.BytesIO)):
        with pytest.raises(ValueError):
            io._restricted_unpickler(_io.BytesIO(payload)).load()


@pytest.mark.gpu
def test_challenge_keynet_imported_from_the_reference_pickle_reproduces_the_notebook():
    import torch
    from PIL import Image
    from keynet_b200 import io, system, torch as ktorch
    z = gu.load('challenge_kat.npz')
    (sensor, knet) = io.load_reference_pickle(PKL)
    assert isinstance(sensor, system.PublicKeyedSensor) and knet.num_parameters() == sum(len(z['layer.%s.W.data' % k]) for k in ['conv1', 'pool1', 'conv2', 'pool2', 'fc1', 'fc2', 'fc3'])
    img = np.array(Image.open(PNG))
    red = img[:, :, 0] if img.ndim == 3 else img
    x = torch.as_tensor(red.astype(np.float32) / 255.0).reshape(1, 1, 28, 28)
    xl = ktorch.affine_to_linear(x)
    assert np.array_equal(xl.numpy(), z['x_linear'])
    y = knet.forward(xl).reshape(-1).cpu().numpy()
    assert np.allclose(y, z['y_printed'], atol=1e-4), y                  # demo/challenge.ipynb cell 5
    assert np.allclose(y, z['y_reference'].reshape(-1)[:-1], atol=2e-5)  # the reference's own forward (partly float64 matrices)
    # a batch goes through the grouped kernels and agrees with the single image
    yb = knet.forward(xl.repeat(64, 1)).reshape(64, -1).cpu().numpy()
    assert np.allclose(yb, y.reshape(1, -1), rtol=1e-4, atol=1e-6)
