# torchrun --nproc-per-node 2 scratch/bench_sharded.py : row-sharded AllConvNet (permutation + gain keys: no shared value blocks),
# NCCL all-gather per layer vs fused epilogue stores to peer memory
import os, sys, time, json, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from keynet_b200 import dist as kdist, system, nets
import bench
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
net = bench.numpy_weights(nets.AllConvNet(), 0).eval()
res = {}
for fused in (False, True):
    np.random.seed(0)
    m = kdist.ShardedKeyedModel((3, 32, 32), net, rank=rank, world=world, fused=fused, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    x = torch.randn(N, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    xc = m.sensor.fromtensor(x.cuda()).encrypt().astensor()
    for _ in range(3):
        y = m.forward_linear(xc)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = m.forward_linear(xc)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res['fused' if fused else 'nccl'] = float(t)
    if fused:
        yp = net(x[:64]).detach().numpy()
        res['max_err_vs_plain'] = float(np.abs(y[:64, :-1].cpu().numpy() - yp).max())
    res['nnz_local_' + ('fused' if fused else 'nccl')] = m.num_parameters_local()
    del m
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({'world': world, 'batch': N, 'ms_per_forward': res}))
dist.barrier(); dist.destroy_process_group()
