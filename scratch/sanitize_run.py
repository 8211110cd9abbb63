"""Small run that touches every kernel family, for compute-sanitizer (memcheck / racecheck):
  python scratch/sanitize_run.py            one GPU: LeNet keynet (CSR, pg_small, pg_cluster), tcgen05 per-pixel and tiled kernels, fused key compile
  torchrun --nproc-per-node 2 scratch/sanitize_run.py dist      two GPUs: fused row-sharded LeNet forward (peer stores + kn_peer_sync)"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from keynet_b200 import nets, system, sparse, engine
from keynet_b200.sparse import MonomialKey

if len(sys.argv) > 1 and sys.argv[1] == 'dist':
    import torch.distributed as dist
    from keynet_b200 import dist as kdist
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
    torch.manual_seed(0)
    net = nets.LeNet_AvgPool().eval()
    np.random.seed(0)
    m = kdist.ShardedKeyedModel((1, 28, 28), net, rank=rank, world=world, fused=True, global_geometric='permutation')
    x = torch.randn(64, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    xc = m.sensor.fromtensor(x.cuda()).encrypt().astensor()
    for _ in range(2):
        y = m.forward(xc).reshape(64, -1).cpu().numpy()
    assert np.allclose(y, net(x).detach().numpy(), atol=1e-4) and not m.sync_timed_out()
    print('[sanitize dist rank %d] ok' % rank)
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0)

torch.manual_seed(0)
net = nets.LeNet_AvgPool().eval()
np.random.seed(0)
(sensor, knet) = system.Keynet((1, 28, 28), net, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
x = torch.randn(128, 1, 28, 28)
for n in (1, 5, 128):
    y = knet.forward(sensor.fromtensor(x[:n].cuda()).encrypt().astensor()).reshape(n, -1).cpu().numpy()
    assert np.allclose(y, net(x[:n]).detach().numpy(), atol=1e-4)
plan = engine.ForwardPlan(sensor, knet, 128, use_graph=False)
plan.run_device(x.cuda())
rs = np.random.RandomState(0)
for (C, M, U, stride) in [(16, 32, 8, 1), (32, 96, 4, 1), (16, 64, 8, 2), (128, 128, 4, 1), (16, 192, 4, 1)]:
    f = rs.randn(M, C, 3, 3).astype(np.float32); b = rs.randn(M).astype(np.float32)
    K = C * U * U + 1
    Ainv = MonomialKey(np.concatenate([rs.permutation(K - 1), [K - 1]]))
    W = sparse.keyed_toeplitz_conv2d((C, U, U), f, b, stride, None, Ainv)
    X = torch.randn(K, 256, device='cuda'); X[-1] = 1
    y = sparse.spmm(W, X, relu=True)
    sparse.tiles_enabled(False); y2 = sparse.spmm(W, X, relu=True); sparse.tiles_enabled(True)
    Wc = sparse.SparseMatrix((W.shape, *W.csr_arrays()))
    y3 = sparse.spmm(Wc, X, relu=True)
    (e1, e2, sc) = (float((y - y3).abs().max()), float((y2 - y3).abs().max()), float(y3.abs().max()))
    print('[sanitize] conv C=%d M=%d U=%d stride=%d tile=%s: max err tiled %.2e, per-pixel %.2e (scale %.1f)' % (C, M, U, stride, W._pg.classes[0].get('tile') is not None, e1, e2, sc), flush=True)
    assert e1 <= 1e-4 * sc and e2 <= 1e-4 * sc
w = torch.randn(300, 9000); bb = torch.randn(300)
A = MonomialKey(np.concatenate([rs.permutation(300), [300]])); Ai = MonomialKey(np.concatenate([rs.permutation(9000), [9000]]))
L = sparse.keyed_linear(w, bb, A, Ai)
y = sparse.spmm(L, torch.randn(9001, 128, device='cuda'))
torch.cuda.synchronize()
print('[sanitize] ok')
