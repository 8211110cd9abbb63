"""VGG16 3x224x224 keyed with a global permutation key on ONE B200 (BASELINE configs[3]/[4] scale: 15.0 G stored
non-zeros, 120 GB as CSR).  Each layer is compiled on the GPU, turned into pattern groups with unique value blocks and
its CSR freed (keep_csr=False), so the whole keyed network occupies < 2 GB.  The keyed forward must equal the plain
network (reference test/test_keynet.py:83-95, atol 1e-3) -- against the pooling the reference actually keys: centred
3x3 / stride 2 / divisor 9 windows (keynet/layer.py:48-56 ignores padding and ceil_mode)."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _he_weights(net, seed):
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for (name, p) in net.named_parameters():
            if p.ndim > 1:
                fan_in = int(np.prod(p.shape[1:]))
                bound = np.sqrt(6.0 / fan_in)           # keeps the activation scale through the ReLUs: logits depend on the input
                p.copy_(torch.from_numpy(rs.uniform(-bound, bound, size=tuple(p.shape)).astype(np.float32)))
            else:
                p.copy_(torch.from_numpy(rs.uniform(-0.1, 0.1, size=tuple(p.shape)).astype(np.float32)))
    return net


def test_vgg16_permutation_keynet_single_gpu():
    from keynet_b200 import system, nets
    free = torch.cuda.mem_get_info()[0]
    if free < 120e9:
        pytest.skip('needs ~120 GB of free HBM for the per-layer compile transients')
    net = _he_weights(nets.VGG16(num_classes=64), 0).eval()
    np.random.seed(0)
    (sensor, knet) = system.Keynet((3, 224, 224), net, global_geometric='permutation', keep_csr=False)
    # closed-form structure: conv k=3,s=1: M*C*(3U-2)^2 + M*U^2 + 1 stored entries (minus dropped zeros)
    names = dict(knet.keyedlayers())
    assert abs(names['conv1_2'].nnz() - (64 * 64 * (3 * 224 - 2) ** 2 + 64 * 224 ** 2 + 1)) < 1000
    assert knet.num_parameters() > 14.9e9
    for (k, L) in knet.keyedlayers():
        if k.startswith('conv'):
            s = L.W._pg.summary()
            assert s['unique_blocks'][0] <= 9, (k, s['unique_blocks'])       # interior, 4 edges, 4 corners
    used = torch.cuda.memory_allocated()
    assert used < 8e9, used
    N = 32
    x = torch.randn(N, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    y = knet.forward(sensor.fromtensor(x.cuda()).encrypt().astensor()).reshape(N, -1).cpu().numpy()
    # plain network with the pooling that is actually keyed
    plain = _he_weights(nets.VGG16(num_classes=64), 0).eval()
    for (k, m) in list(plain.named_children()):
        if isinstance(m, nn.AvgPool2d):
            setattr(plain, k, nn.AvgPool2d(3, 2, 1, ceil_mode=False, count_include_pad=True))
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        yp = plain.cuda()(x.cuda()).cpu().numpy()
    scale = np.abs(yp).max()
    assert scale > 0.05 and np.std(yp, axis=0).max() > 1e-3 * scale             # logits are input dependent
    assert np.allclose(y, yp, atol=1e-3 * max(1.0, scale)), (np.abs(y - yp).max(), scale)
    assert np.array_equal(y.argmax(1), yp.argmax(1))
    # the benchmarked batch (256: tcgen05 kernels, two batch tiles) through the engine the bench uses
    from keynet_b200 import engine
    N = 256
    plan = engine.ForwardPlan(sensor, knet, N, use_graph=False)
    x = torch.randn(N, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    y = plan.run_device(x.cuda()).cpu().numpy()
    with torch.no_grad():
        yp = torch.cat([plain(x[i:i + 64].cuda()) for i in range(0, N, 64)]).cpu().numpy()
    scale = np.abs(yp).max()
    assert np.allclose(y, yp, atol=1e-3 * max(1.0, scale)), (np.abs(y - yp).max(), scale)
    assert np.array_equal(y.argmax(1), yp.argmax(1))
