# where does the keyed VGG16 drift from the plain network?  conv+relu outputs of a PermutationKeynet carry the identity key,
# so they compare directly with the plain network's activations (torch fp32 on the GPU, and fp64 on the CPU for 2 images)
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from keynet_b200 import system, sparse, layer as klayer
wl = bench.workload('vgg16')
net = wl['net']
np.random.seed(0)
(sensor, knet) = system.Keynet(wl['inshape'], net, **wl['keys'])
N = 2
x = torch.randn((N,) + wl['inshape'], generator=torch.Generator().manual_seed(1))
acts = {}
net64 = __import__('copy').deepcopy(net).double()
hooks = [m.register_forward_hook(lambda mod, i, o, k=k: acts.__setitem__(k, o.detach())) for (k, m) in net64.named_children()]
with torch.no_grad():
    net64(x.double())
h = sensor.fromtensor(x.cuda()).encrypt().astensor()
names = [k for (k, m) in knet._keynet.named_children()]
for (k, m) in knet._keynet.named_children():
    h = m.forward(h)
    if isinstance(m, klayer.KeyedLayer) and k.startswith('conv'):
        rk = k.replace('conv', 'relu')
        ref = acts[rk].reshape(N, -1).numpy()
        got = h[:, :-1].cpu().numpy().astype(np.float64)
        print('%-8s max|ref| %.4f  max abs err %.3e  rel(max err / max ref) %.2e  hom err %.1e' % (k, np.abs(ref).max(), np.abs(got - ref).max(), np.abs(got - ref).max() / np.abs(ref).max(), float((h[:, -1] - 1).abs().max())))
    elif isinstance(m, klayer.KeyedLayer):
        print('%-8s (keyed output) max %.4f hom err %.1e' % (k, float(h[:, :-1].abs().max()), float((h[:, -1] - 1).abs().max())))
