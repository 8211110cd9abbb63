"""Parity AT THE BENCHMARKED SHAPES: the kernels bench.py times -- pg_tc_kernel (tcgen05, NB=2, DUAL,
2-CTA multicast, block_of dedup, ragged group_k) on AllConvNet at batch 256 / 4096, the non-DUAL multi-chunk instantiation on
a full-size VGG16 conv3_2 -- are compared with the oracle (csr_matvecs on the same compiled CSR, or on an oracle-keyed row
band for the layer no host can hold), with the plain network, and through the same ForwardPlan buffers the bench reads.
Bars: rtol 1e-4, atol 1e-5*max|y|; argmax equal to the plain net."""
import numpy as np
import pytest
import torch
from torch import nn

import bench

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


@pytest.fixture(scope='module')
def acn():
    from keynet_b200 import system
    wl = bench.workload('acn')
    np.random.seed(0)
    (sensor, knet) = system.Keynet(wl['inshape'], wl['net'], **wl['keys'])
    return (wl, sensor, knet, bench.oracle_layers_from_gpu(sensor, knet))


@pytest.mark.parametrize('N', [256, 4096])
def test_allconvnet_forwardplan_at_bench_batch_matches_oracle_and_plain_net(acn, N):
    from keynet_b200 import engine
    (wl, sensor, knet, layers) = acn
    plan = engine.ForwardPlan(sensor, knet, N, use_graph=False)
    images = torch.randn((N,) + wl['inshape'], device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
    plan.run_device(images)
    torch.cuda.synchronize()
    # the tensor-core kernel really is the one that ran for the dominant layers
    conv2 = dict(knet.keyedlayers())['conv2'].W
    assert conv2._pg is not None and any(c['tc'] is not None for c in conv2._pg.classes)
    chk = bench.check_plan_against_oracle(plan, layers, n=64)
    assert chk['ok'], chk
    p = bench.check_against_plain_net(wl, images, plan.logits, n=64)
    assert p['argmax_equal'] and p['max_abs_err_vs_plain_net'] <= 1e-4 * max(1.0, p['max_abs_logit']), p
    # the LAST images of the batch too (upper batch tiles / super-tiles of the rasterised grid)
    from oracle import keynet_oracle as ko
    x = ko.affine_to_linear(images[-32:].cpu().numpy())
    ref = ko.linear_to_affine(ko.keyed_forward(layers, x, threads=bench.host_threads()))
    (bad, rel) = bench._close_frac(plan.logits[-32:].cpu().numpy(), ref)
    assert bad == 0.0, (bad, rel)


class _OneConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)

    def forward(self, x):
        return self.conv1(x)


@pytest.mark.parametrize('shape', [(256, 256, 56), (64, 64, 112)])
def test_full_size_vgg_conv_layer_matches_oracle_band(shape):
    """VGG16 conv3_2 (G = 256, K = 2305: row chunks, non-DUAL) and a conv1_2-like layer (G = 64: DUAL) at full channel count,
    permuted input key, batch 256, CSR dropped: the rows of an oracle-keyed band against the timed output buffer."""
    from keynet_b200 import system, engine
    (cin, cout, U) = shape
    if torch.cuda.mem_get_info()[0] < 100e9:
        pytest.skip('needs ~60 GB of free HBM for the compile transients')
    net = bench.he_weights(_OneConv(cin, cout), 1).eval()
    wl = dict(name='oneconv', net=net, inshape=(cin, U, U), keys=dict(global_geometric='permutation'), label='one conv')
    np.random.seed(0)
    (sensor, knet) = system.Keynet(wl['inshape'], net, keep_csr=False, **wl['keys'])
    N = 256
    plan = engine.ForwardPlan(sensor, knet, N, use_graph=False)
    images = torch.randn((N,) + wl['inshape'], device='cuda', generator=torch.Generator(device='cuda').manual_seed(5))
    plan.run_device(images)
    torch.cuda.synchronize()
    W = dict(knet.keyedlayers())['conv1'].W
    assert W._pg is not None and any(c['tc'] is not None for c in W._pg.classes)
    bands = bench.oracle_bands_on_cpu(wl, 1.0 / 400.0)
    chk = bench.check_plan_against_bands(plan, bands, n=8)
    assert chk['ok'], chk
    del plan, knet, sensor
    torch.cuda.empty_cache()
