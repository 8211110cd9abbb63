"""Fused conv (+ReLU) -> average pooling (csrc/convpool.cu, engine.ForwardPlan): same logits as the unfused chain and as the
oracle's csr_matvecs chain, for permutation keys on LeNet (both conv/pool pairs fuse) and on odd batch sizes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('keys', [dict(global_geometric='permutation'), dict(), dict(global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1))])
@pytest.mark.parametrize('N', [4, 36, 128, 1000])
def test_fused_lenet_matches_unfused_and_oracle(keys, N):
    from keynet_b200 import system, nets, engine
    from oracle import keynet_oracle as ko
    import bench
    torch.manual_seed(1)
    net = nets.LeNet_AvgPool().eval()
    np.random.seed(2)
    (sensor, knet) = system.Keynet((1, 28, 28), net, **keys)
    x = torch.randn(N, 1, 28, 28, generator=torch.Generator().manual_seed(N)).cuda()
    try:
        engine.fusion_enabled(True)
        plan = engine.ForwardPlan(sensor, knet, N, use_graph=False)
    finally:
        engine.fusion_enabled(False)
    assert sorted(plan.fused) == [1, 3]                      # conv1+pool1, conv2+pool2
    y = plan.run_device(x).clone()
    plain = engine.ForwardPlan(sensor, knet, N, use_graph=False)
    assert plain.fused == {}
    y0 = plain.run_device(x).clone()
    assert torch.allclose(y, y0, rtol=1e-4, atol=1e-6 * float(y0.abs().max()) + 1e-7)
    layers = bench.oracle_layers_from_gpu(sensor, knet)
    ref = ko.linear_to_affine(ko.keyed_forward(layers, ko.affine_to_linear(x.cpu().numpy()), threads=4))
    (bad, rel) = bench._close_frac(y.cpu().numpy(), ref)
    assert bad == 0.0, (bad, rel)
    assert np.allclose(y.cpu().numpy(), net(x.cpu()).detach().numpy(), atol=1e-4)


def test_gain_keys_are_not_fused():
    from keynet_b200 import system, nets, engine
    torch.manual_seed(1)
    net = nets.LeNet_AvgPool().eval()
    np.random.seed(2)
    (sensor, knet) = system.Keynet((1, 28, 28), net, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    try:
        engine.fusion_enabled(True)
        assert engine.ForwardPlan(sensor, knet, 64, use_graph=False).fused == {}
    finally:
        engine.fusion_enabled(False)
