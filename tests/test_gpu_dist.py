"""Row-sharded keyed model on the GPU (single rank = world 1 exercises the shard-major compile + forward; with 2+
visible GPUs a 2-rank NCCL run is launched through torch.multiprocessing)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net():
    from keynet_b200 import nets
    torch.manual_seed(0)
    return nets.LeNet_AvgPool().eval()


def _reference(x):
    from keynet_b200 import system
    np.random.seed(0)
    (sensor, knet) = system.Keynet((1, 28, 28), _net(), global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    xc = sensor.fromtensor(x).encrypt().astensor()
    return (xc, knet.forward(xc).reshape(x.shape[0], -1).numpy())


def test_sharded_model_world1_matches_unsharded():
    from keynet_b200 import dist as kdist
    x = torch.randn(64, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    (xc, y_ref) = _reference(x)
    np.random.seed(0)
    m = kdist.ShardedKeyedModel((1, 28, 28), _net(), rank=0, world=1, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    y = m.forward(xc).reshape(64, -1).cpu().numpy()
    assert np.allclose(y, y_ref, rtol=1e-4, atol=1e-5)
    assert np.allclose(y, _net()(x).detach().numpy(), atol=1e-4)


def _worker(rank, world, port, q, fused=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from keynet_b200 import dist as kdist
        x = torch.randn(64, 1, 28, 28, generator=torch.Generator().manual_seed(1))
        (xc, y_ref) = _reference(x)
        np.random.seed(0)
        m = kdist.ShardedKeyedModel((1, 28, 28), _net(), rank=rank, world=world, fused=fused, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
        y = m.forward(xc).reshape(64, -1).cpu().numpy()
        y = m.forward(xc).reshape(64, -1).cpu().numpy()       # twice: ping-pong buffers are reused
        q.put((rank, bool(np.allclose(y, y_ref, rtol=1e-4, atol=1e-5)), float(np.abs(y - y_ref).max()), m.num_parameters_local()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('fused', [False, True])
def test_sharded_model_world2(fused):
    """fused=False: NCCL all-gather per layer; fused=True: epilogue stores to peer memory (K5), no NCCL on the data path."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, fused)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for (rank, ok, err, nnz) in res:
        assert ok, (rank, err)
    assert abs(res[0][3] - res[1][3]) < 0.2 * max(res[0][3], res[1][3])      # shards are balanced
