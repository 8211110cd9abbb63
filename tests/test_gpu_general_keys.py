"""General (non-monomial) keys on the GPU: the SpGEMM of csrc/spgemm.cu against the oracle's csr_matmat restatement, and
networks keyed with Givens-orthogonal / doubly stochastic / affine keys against the reference fixture and the plain net."""
import numpy as np
import pytest
import torch
from torch import nn

from tests import golden_util as gu
from tests.test_general_keys import CFG, _golden_net

pytestmark = pytest.mark.gpu


def _rand_csr(rs, n_rows, n_cols, density, long_row=None):
    from oracle import keynet_oracle as ko
    D = (rs.rand(n_rows, n_cols) < density) * rs.randn(n_rows, n_cols)
    if long_row is not None:
        D[long_row, :] = rs.randn(n_cols)
    D[min(3, n_rows - 1), :] = 0                    # an empty row
    return ko.csr_from_dense(D.astype(np.float32))


@pytest.mark.parametrize('shape', [(40, 50, 60, 0.2), (300, 200, 9000, 0.02), (64, 64, 64, 1.0)])
def test_spgemm_matches_oracle_csr_matmat(shape):
    from keynet_b200.sparse import SparseMatrix, _spgemm_device
    from oracle import keynet_oracle as ko
    (m, k, n, dens) = shape
    rs = np.random.RandomState(m)
    A = _rand_csr(rs, m, k, dens, long_row=0)
    B = _rand_csr(rs, k, n, dens if n < 1000 else 0.05, long_row=1)       # the wide case gives rows of > 8192 products (global sort path)
    if m == 64:                                                            # exact cancellation: zeros must be dropped
        A = ko.csr_from_dense(np.kron(np.eye(32), np.array([[1, 1], [1, -1]])).astype(np.float32))
        B = ko.csr_from_dense(np.kron(np.eye(32), np.array([[2, 3], [2, 5]])).astype(np.float32))
    ref = ko.sort_indices(ko.matmat(A, B))
    (SA, SB) = (SparseMatrix((A.shape, A.indptr, A.indices, A.data)), SparseMatrix((B.shape, B.indptr, B.indices, B.data)))
    (ip, ix, dt) = _spgemm_device((SA._indptr, SA._indices, SA._data), A.shape[0], (SB._indptr, SB._indices, SB._data))
    assert np.array_equal(ip.cpu().numpy(), ref.indptr) and np.array_equal(ix.cpu().numpy(), ref.indices)
    assert np.allclose(dt.cpu().numpy(), ref.data, rtol=1e-5, atol=1e-5)
    if m == 64:
        assert np.array_equal(dt.cpu().numpy(), ref.data)


def test_lenet_givens_affine_keynet_matches_reference():
    """The reference's LeNet orthogonal configuration (test/test_keynet.py:180-197): same seed -> same keys -> same
    compiled structure (indices bit-exact, values to fp32 rounding), forward equal to the reference's and the plain net's."""
    from keynet_b200 import system
    from keynet_b200.sparse import SparseKey
    z = gu.load('lenet_givens.npz')
    net = _golden_net(z)
    np.random.seed(0)
    (sensor, knet) = system.Keynet((1, 28, 28), net, **CFG)
    assert isinstance(sensor.keypair()[0], SparseKey)
    layers = gu.jstr(z, 'layers')
    assert [k for (k, _) in knet.keyedlayers()] == layers
    for (k, L) in knet.keyedlayers():
        (shape, indptr, indices, data) = gu.csr_arrays(z, 'layer.%s.W' % k)
        from oracle import keynet_oracle as ko
        ref = ko.sort_indices(ko.csr(shape, indptr, indices, data.astype(np.float32)))
        (ip, ix, dt) = L.W.csr_arrays()
        assert tuple(L.W.shape) == tuple(shape), k
        assert np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices), k
        assert np.allclose(dt, ref.data, rtol=1e-5, atol=2e-6), k
    assert knet.num_parameters() == int(z['num_parameters'])
    x = torch.from_numpy(z['x'])
    xc = sensor.fromtensor(x).encrypt().astensor()
    assert np.allclose(xc.numpy(), z['x_cipher'], rtol=1e-4, atol=1e-5)
    y = knet.forward(xc).reshape(x.shape[0], -1).numpy()
    assert np.allclose(y, z['logits_keyed'], rtol=1e-4, atol=1e-4)
    assert np.allclose(y, z['logits_plain'], atol=1e-4)
    assert np.array_equal(y.argmax(1), z['logits_plain'].argmax(1))
    # decrypt(encrypt(x)) == x
    xd = sensor.fromtensor(x).encrypt().decrypt().astensor()
    assert np.allclose(xd.numpy(), x.numpy(), atol=1e-4)


class _Tiny(nn.Module):
    def __init__(self):
        super(_Tiny, self).__init__()
        self.conv1 = nn.Conv2d(2, 4, 3, padding=1); self.relu1 = nn.ReLU()
        self.pool1 = nn.AvgPool2d(3, 2, 1)
        self.conv2 = nn.Conv2d(4, 6, 3, padding=1); self.relu2 = nn.ReLU()
        self.fc1 = nn.Linear(6 * 4 * 4, 7)

    def forward(self, x):
        x = self.relu2(self.conv2(self.pool1(self.relu1(self.conv1(x)))))
        return self.fc1(x.reshape(x.shape[0], -1))


@pytest.mark.parametrize('kwargs', [
    dict(local_geometric='doubly_stochastic', alpha=2.0, blocksize=4, local_photometric='uniform_random_affine', beta=1.0, gamma=1.0),
    dict(global_geometric='givens_orthogonal', alpha=40, global_photometric='uniform_random_gain', beta=1.0),
    dict(local_geometric='givens_orthogonal', alpha=8, blocksize=4, global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2),
         hierarchical_permute_at_level=(0,), memoryorder='block'),
])
def test_general_keys_keyed_net_equals_plain_net(kwargs):
    from keynet_b200 import system
    torch.manual_seed(0)
    net = _Tiny().eval()
    np.random.seed(1)
    (sensor, knet) = system.Keynet((2, 8, 8), net, **kwargs)
    x = torch.randn(64, 2, 8, 8, generator=torch.Generator().manual_seed(2))
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(64, -1).numpy()
    yp = net(x).detach().numpy()
    assert np.allclose(y, yp, atol=2e-3 if 'doubly_stochastic' in str(kwargs) else 2e-4), np.abs(y - yp).max()
    assert np.array_equal(y.argmax(1), yp.argmax(1))


def test_tiled_orthogonal_keynet_factory():
    """TiledOrthogonalKeynet (keynet/system.py:504-510) on LeNet: keyed == plain."""
    from keynet_b200 import system, nets
    torch.manual_seed(0)
    net = nets.LeNet_AvgPool().eval()
    np.random.seed(0)
    (sensor, knet) = system.TiledOrthogonalKeynet((1, 28, 28), net, 4)
    x = torch.randn(8, 1, 28, 28, generator=torch.Generator().manual_seed(3))
    y = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(8, -1).numpy()
    yp = net(x).detach().numpy()
    assert np.allclose(y, yp, atol=1e-3), np.abs(y - yp).max()


def test_sensor_image_roundtrip_through_png(tmp_path):
    """KeyedSensor.load -> encrypt -> save (8-bit PNG of the encrypted image + image key) -> load(imagekey) recovers the
    image to quantisation (keynet/system.py:173-207, README quickstart with demo/owl.jpg)."""
    from PIL import Image
    from keynet_b200 import system
    rs = np.random.RandomState(0)
    src = tmp_path / 'in.png'
    Image.fromarray(rs.randint(0, 255, size=(28, 28)).astype(np.uint8)).save(src)
    np.random.seed(0)
    (sensor, _) = system.Keynet((1, 28, 28), None, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    x = sensor.load(str(src)).tensor().clone()
    assert x.shape == (1, 1, 28, 28) and float(x.max()) <= 255
    (outfile, imagekey) = sensor.encrypt().save(str(tmp_path / 'enc.png'))
    enc = np.asarray(Image.open(outfile))
    assert enc.shape == (28, 28) and not np.array_equal(enc, np.asarray(Image.open(src)))
    assert sensor.asimage().size == (28, 28)
    y = sensor.load(outfile, imagekey).tensor()
    step = float(sensor_range(x, sensor)) / 255.0
    assert np.abs(y.numpy() - x.numpy()).max() <= 2.0 * step + 1e-3


def sensor_range(x, sensor):
    """Dynamic range of the encrypted image (gain keys in [1, 2)): quantisation step of the PNG, seen after decryption."""
    return 2.0 * float(x.max() - x.min()) + 1.0


def test_layer_spy_picture():
    from keynet_b200 import system, nets
    torch.manual_seed(0)
    np.random.seed(0)
    (sensor, knet) = system.PermutationKeynet((1, 28, 28), nets.LeNet_AvgPool().eval())
    L = dict(knet.keyedlayers())['conv1']
    im = L.spy(mindim=128, showdim=256)
    assert im.mode == 'RGB' and max(im.size) == 256
