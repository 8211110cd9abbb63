"""Tiled keynets WITHOUT the expanded matrices (BASELINE configs[3] as written: "VGG16 tiled keynet, unique tiles"):
Conv2dTiledMatrix.from_twin takes the tile tables from a one-channel twin of the layer instead of the 120 GB CSR.
Checks: (i) equal to the tiled view of the expanded matrix wherever that fits; (ii) VGG16 TiledIdentityKeynet(56) /
TiledPermutationKeynet(14) (reference test/test_keynet.py:98-130) -- stored-parameter counts of SURVEY.md 8d (351.7 M /
201.2 M floats in the conv layers) and keyed == plain network (atol 1e-3)."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tile', [4, 8])
@pytest.mark.parametrize('local', ['identity', 'permutation'])
def test_conv_tiled_from_twin_equals_tiled_view_of_expanded_matrix(tile, local):
    from keynet_b200 import system, layer
    torch.manual_seed(0)
    m = nn.Conv2d(8, 12, 3, padding=1)
    (inshape, outshape) = ((8, 16, 16), (12, 16, 16))
    np.random.seed(3)
    kw = dict(blocksize=tile, tileshape=(tile, tile)) if local != 'identity' else dict(tileshape=(tile, tile))
    (A, _) = system.keygen(outshape, 'identity', local, 'identity', 'identity', **kw)
    (_, Ainv) = system.keygen(inshape, 'identity', local, 'identity', 'identity', **kw)
    full = layer.KeyedLayer(m, inshape, outshape, A, Ainv, tileshape=(tile, tile), keep_csr=True)
    twin = layer.KeyedLayer(m, inshape, outshape, A, Ainv, tileshape=(tile, tile), keep_csr=False)
    assert full.W._data is not None and twin.W._data is None
    assert twin.W.blocks() == full.W.blocks()
    assert twin.W.nnz() == full.W.nnz() and twin.W._n_tile_entries == full.W._n_tile_entries
    assert twin.W.expanded_nnz() == full.W.expanded_nnz()
    x = torch.randn(64, 8 * 256 + 1, device='cuda'); x[:, -1] = 1
    assert torch.allclose(twin.forward(x), full.forward(x), rtol=1e-4, atol=1e-5)


def _he(net, seed):
    import bench
    return bench.he_weights(net, seed)


@pytest.mark.slow
@pytest.mark.parametrize('factory,tilesize,conv_floats', [('TiledIdentityKeynet', 56, 351.8e6), ('TiledIdentityKeynet', 14, 201.2e6), ('TiledPermutationKeynet', 14, None)])
def test_vgg16_tiled_keynet_parameter_count_and_forward(factory, tilesize, conv_floats):
    from keynet_b200 import system, nets, tiled
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip('needs ~40 GB of free HBM')
    net = _he(nets.VGG16(num_classes=64), 0).eval()
    np.random.seed(0)
    (sensor, knet) = getattr(system, factory)((3, 224, 224), net, tilesize, keep_csr=False)
    conv = sum(L.W._n_spatial_entries * L._outshape[0] * L._inshape[0] for (k, L) in knet.keyedlayers() if isinstance(L.W, tiled.Conv2dTiledMatrix))
    if conv_floats is not None:
        # a probe with the reference's plain TiledMatrix on the (0,0) channel block (entries * Cout * Cin);
        # Conv2dTiledMatrix itself also stores a (0,0) element for every unique tile (sparse.py:693-717), a few entries more
        assert abs(conv - conv_floats) < 0.03 * conv_floats, conv
    conv_floats = conv
    fc = sum(L.nnz() for (k, L) in knet.keyedlayers() if k.startswith('fc'))
    fc_expected = (25088 * 4096 + 4096 + 1) + (4096 * 4096 + 4096 + 1) + (4096 * 64 + 64 + 1)      # fc6-8 stay CSR (num_classes=64 here; 130.3 M with 2622)
    assert abs(fc - fc_expected) < 1000
    assert knet.num_parameters() < 1.45 * (conv_floats + fc_expected)               # ~100x smaller than the 15.0 G entries of the expansion
    assert torch.cuda.memory_allocated() < 12e9
    N = 32
    x = torch.randn(N, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    y = knet.forward(sensor.fromtensor(x.cuda()).encrypt().astensor()).reshape(N, -1).cpu().numpy()
    plain = _he(nets.VGG16(num_classes=64), 0).eval()
    for (k, m) in list(plain.named_children()):
        if isinstance(m, nn.AvgPool2d):
            setattr(plain, k, nn.AvgPool2d(3, 2, 1, ceil_mode=False, count_include_pad=True))
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        yp = plain.cuda()(x.cuda()).cpu().numpy()
    scale = np.abs(yp).max()
    assert np.allclose(y, yp, atol=1e-3 * max(1.0, scale)), (np.abs(y - yp).max(), scale)
    assert np.array_equal(y.argmax(1), yp.argmax(1))
