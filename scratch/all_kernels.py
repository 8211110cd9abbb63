# exercise every kernel once at a representative size (for one ncu pass): compile + forward of AllConvNet conv layer, LeNet at batch 65536
import sys, numpy as np, torch
sys.path.insert(0, '.')
from keynet_b200 import system, nets, engine
import bench
torch.manual_seed(0)
net = bench.numpy_weights(nets.LeNet_AvgPool(), 0).eval()
np.random.seed(0)
(sensor, knet) = system.Keynet((1, 28, 28), net, global_geometric='permutation', global_photometric='uniform_random_affine', beta=1.0, gamma=1.0)
plan = engine.ForwardPlan(sensor, knet, 65536, use_graph=False)
plan.images.copy_(torch.randn(65536, 784, device='cuda'))
for _ in range(2):
    plan.run_device()
torch.cuda.synchronize()
wl = bench.workload('acn')
np.random.seed(0)
(s2, k2) = system.Keynet(wl['inshape'], wl['net'], **wl['keys'])
torch.cuda.synchronize()
print('done')
