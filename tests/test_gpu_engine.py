"""ForwardPlan (keynet_b200/engine.py): the batched encrypt + forward chain with pre-allocated activations, eager, as a CUDA
graph, through host buffers, and software-pipelined over several host batches -- all equal to the layer-by-layer API."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _keyed_lenet(**kw):
    from keynet_b200 import system, nets
    torch.manual_seed(0)
    net = nets.LeNet_AvgPool().eval()
    np.random.seed(0)
    (sensor, knet) = system.Keynet((1, 28, 28), net, **kw)
    return (net, sensor, knet)


@pytest.mark.parametrize('kw', [dict(global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0),
                                dict(global_geometric='permutation', global_photometric='uniform_random_affine', beta=1.0, gamma=1.0),
                                dict(local_geometric='givens_orthogonal', alpha=2.0, blocksize=7)])
@pytest.mark.parametrize('use_graph', [False, True])
def test_forward_plan_equals_layer_api(kw, use_graph):
    from keynet_b200 import engine
    (net, sensor, knet) = _keyed_lenet(**kw)
    N = 256
    x = torch.randn(N, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    ref = knet.forward(sensor.fromtensor(x).encrypt().astensor()).reshape(N, -1).numpy()
    plan = engine.ForwardPlan(sensor, knet, N, use_graph=use_graph)
    y = plan.run_device(x.cuda()).cpu().numpy()
    assert np.allclose(y, ref, rtol=1e-5, atol=1e-6)
    assert np.allclose(y, net(x).detach().numpy(), atol=2e-4)
    # host buffers in / out
    yh = plan.run_host(x.pin_memory()).numpy()
    assert np.array_equal(yh, y)


def test_pipelined_host_batches_equal_one_by_one():
    from keynet_b200 import engine
    (net, sensor, knet) = _keyed_lenet(global_geometric='permutation')
    N = 512
    plan = engine.ForwardPlan(sensor, knet, N, use_graph=False)
    g = torch.Generator().manual_seed(2)
    batches = [torch.randn(N, 1, 28, 28, generator=g).pin_memory() for _ in range(5)]
    one_by_one = [plan.run_host(b).numpy().copy() for b in batches]
    outs = [torch.empty((N, plan.K), dtype=torch.float32).pin_memory() for _ in range(5)]
    plan.run_host_many(batches, outs)
    for (a, b) in zip(one_by_one, outs):
        assert np.array_equal(a, b.numpy())
    # reusing two output / input buffers round-robin (what bench.py does) is safe for the LAST batches
    outs2 = [torch.empty((N, plan.K), dtype=torch.float32).pin_memory() for _ in range(2)]
    plan.run_host_many(batches[:4], [outs2[k % 2] for k in range(4)])
    assert np.array_equal(outs2[0].numpy(), one_by_one[2]) and np.array_equal(outs2[1].numpy(), one_by_one[3])
