# VGG16 keyed forward: tensor-core (3xTF32) path vs fp32-FMA path vs the plain torch network (fp32 on CPU / fp64 on CPU)
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from keynet_b200 import system, sparse
wl = bench.workload('vgg16')
np.random.seed(0)
(sensor, knet) = system.Keynet(wl['inshape'], wl['net'], **wl['keys'])
N = 128
x = torch.randn((N,) + wl['inshape'], generator=torch.Generator().manual_seed(1))
xc = sensor.fromtensor(x.cuda()).encrypt().astensor()
out = {}
for tc in (True, False):
    sparse.tensor_cores_enabled(tc)
    y = knet.forward(xc).reshape(N, -1).cpu().numpy().astype(np.float64)
    out[tc] = y
with torch.no_grad():
    y32 = wl['net'](x[:8]).numpy().astype(np.float64)
    y64 = wl['net'].double()(x[:8].double()).numpy()
print('logit scale', np.abs(y64).max(), 'spread across classes', y64.std())
print('plain fp32 vs fp64      ', np.abs(y32 - y64).max())
print('keyed TC   vs plain fp64', np.abs(out[True][:8] - y64).max())
print('keyed FMA  vs plain fp64', np.abs(out[False][:8] - y64).max())
print('keyed TC   vs keyed FMA ', np.abs(out[True] - out[False]).max())
print('argmax TC/FMA/plain equal', np.array_equal(out[True][:8].argmax(1), y64.argmax(1)), np.array_equal(out[False][:8].argmax(1), y64.argmax(1)))
