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


def _worker(rank, world, port, q, fused=False, selective=True):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from keynet_b200 import dist as kdist
        x = torch.randn(64, 1, 28, 28, generator=torch.Generator().manual_seed(1))
        (xc, y_ref) = _reference(x)
        np.random.seed(0)
        m = kdist.ShardedKeyedModel((1, 28, 28), _net(), rank=rank, world=world, fused=fused, selective=selective, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
        if fused and selective:                                # rows no peer asked for must never be read: poison the buffers
            for (t, h) in m._symm_buffers(64, torch.device('cuda', rank)):
                t.fill_(float('nan'))
            dist.barrier()
        y = m.forward(xc).reshape(64, -1).cpu().numpy()
        y = m.forward(xc).reshape(64, -1).cpu().numpy()       # twice: ping-pong buffers are reused
        ok = bool(np.allclose(y, y_ref, rtol=1e-4, atol=1e-5)) and not m.sync_timed_out()
        if fused and selective:                                # the same chain replayed as CUDA graphs (one per ping-pong parity)
            m.capture(64)
            X = xc.cuda().t().contiguous()
            for _ in range(3):
                yg = m.forward_graph(X)[:, :-1].cpu().numpy()
                ok = ok and bool(np.allclose(yg, y_ref, rtol=1e-4, atol=1e-5))
            ok = ok and not m.sync_timed_out()
        q.put((rank, ok, float(np.abs(y - y_ref).max()), m.num_parameters_local(), m.peer_store_fraction()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('fused,selective', [(False, False), (True, False), (True, True)])
def test_sharded_model_world2(fused, selective):
    """fused=False: NCCL all-gather per layer; fused=True: epilogue stores to peer memory (K5), no NCCL on the data path;
    selective: only the rows a peer's next layer reads are stored to it (halo exchange instead of all-gather)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, fused, selective)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for (rank, ok, err, nnz, frac) in res:
        assert ok, (rank, err)
        if fused and selective:
            assert frac is not None and frac < 0.95, frac          # fewer stores than a full all-gather
    assert abs(res[0][3] - res[1][3]) < 0.2 * max(res[0][3], res[1][3])      # shards are balanced


def test_sharded_model_world1_general_keys():
    """Row-sharded compile of general-key layers (Givens + affine keys): full compile, then row gather + column remap."""
    from keynet_b200 import dist as kdist, system
    kw = dict(local_geometric='givens_orthogonal', alpha=2.0, blocksize=7, local_photometric='uniform_random_affine', beta=1.0, gamma=1.0)
    x = torch.randn(64, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    np.random.seed(0)
    (sensor, knet) = system.Keynet((1, 28, 28), _net(), **kw)
    xc = sensor.fromtensor(x).encrypt().astensor()
    y_ref = knet.forward(xc).reshape(64, -1).numpy()
    np.random.seed(0)
    m = kdist.ShardedKeyedModel((1, 28, 28), _net(), rank=0, world=1, **kw)
    y = m.forward(xc).reshape(64, -1).cpu().numpy()
    assert np.allclose(y, y_ref, rtol=1e-4, atol=1e-5)
    assert np.allclose(y, _net()(x).detach().numpy(), atol=2e-4)


@pytest.mark.parametrize('photometric', ['uniform_random_bias', 'uniform_random_affine', 'uniform_random_gain'])
def test_row_shard_with_bias_keys_equals_rows_of_the_full_compile(photometric):
    """a row shard compiled with a gathered column layout (col_remap) must carry the input key's bias
    column.  One GPU is enough: shard rows + remapped columns of W_hat against the same rows of the unsharded compile."""
    from keynet_b200 import system, sparse
    rs = np.random.RandomState(0)
    np.random.seed(1)
    (A, _) = system.keygen((16, 14, 14), 'permutation', 'identity', photometric, 'identity', beta=1.0, gamma=1.0)
    (_, Ainv) = system.keygen((6, 14, 14), 'permutation', 'identity', photometric, 'identity', beta=1.0, gamma=1.0)
    f = rs.randn(16, 6, 3, 3).astype(np.float32); b = rs.randn(16).astype(np.float32)
    full = sparse.keyed_toeplitz_conv2d((6, 14, 14), f, b, 1, A, Ainv, build_groups=False)
    (R, K) = full.shape
    rows = np.sort(rs.permutation(R - 1)[:500])
    n_phys = K + 37
    remap = np.concatenate([rs.permutation(n_phys - 1)[:K - 1], [n_phys - 1]]).astype(np.int64)       # homogeneous coordinate stays last
    shard = sparse.keyed_toeplitz_conv2d((6, 14, 14), f, b, 1, A, Ainv, rows=rows, build_groups=False, col_remap=remap, n_cols_phys=n_phys)
    assert shard.shape == (len(rows), n_phys)
    D = full.todense()[rows]
    S = shard.todense()
    E = np.zeros_like(S)
    E[:, remap] = D
    assert np.array_equal(E.view(np.uint32), S.view(np.uint32))
    if photometric != 'uniform_random_gain':
        assert np.count_nonzero(S[:, -1]) > 0.9 * len(rows)            # the bias column is really there


def test_shard_save_and_load_roundtrip(tmp_path):
    """every rank saves / loads its own shard (canonical CSR over the gathered column layout + shard bookkeeping)."""
    from keynet_b200 import dist as kdist, io
    x = torch.randn(48, 1, 28, 28, generator=torch.Generator().manual_seed(2))
    (xc, y_ref) = _reference(x)
    np.random.seed(0)
    m = kdist.ShardedKeyedModel((1, 28, 28), _net(), rank=0, world=1, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    p = io.save_shard(str(tmp_path / 'rank0.keynet'), m)
    m2 = io.load_shard(p, rank=0, world=1)
    assert m2.num_parameters_local() == m.num_parameters_local()
    for (a, b) in zip(m.layers, m2.layers):
        (ia, xa, da) = a.W.csr_arrays(); (ib, xb, db) = b.W.csr_arrays()
        assert np.array_equal(ia - ia[0], ib) and np.array_equal(xa[ia[0]:ia[-1]], xb) and np.array_equal(da[ia[0]:ia[-1]].view(np.uint32), db.view(np.uint32))
    y = m2.forward(xc).reshape(48, -1).cpu().numpy()
    assert np.allclose(y, y_ref, rtol=1e-4, atol=1e-5)
    with pytest.raises(AssertionError):
        io.load_shard(p, rank=1, world=2)
