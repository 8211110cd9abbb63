import sys, numpy as np, torch, warnings
sys.path.insert(0, '.')
from keynet_b200 import system, nets
import bench
net = bench.numpy_weights(nets.AllConvNet(), 0).eval()
np.random.seed(0)
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter('always')
    (sensor, knet) = system.Keynet((3, 32, 32), net, global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    for x in w: print('WARN', x.message)
for (k, L) in knet.keyedlayers():
    print(k, L.W._pg.summary() if L.W._pg is not None else None)
