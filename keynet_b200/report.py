"""Parameter-count table of the paper (reference: demo/figures.py:236-293 `print_parameters`): stored parameters of the plain
network and of its keyed versions -- IdentityKeynet, PermutationKeynet, TiledPermutationKeynet-k, TiledOrthogonalKeynet-k --
for LeNet, AllConvNet and VGG-16.  `num_parameters()` of a keyed network = stored entries of its layer matrices (CSR entries,
or elements of the unique tiles for tiled layers, keynet/system.py:152-154, keynet/sparse.py:649).

B200 specifics: VGG-16 / AllConvNet rows are built with keep_csr=False (pattern groups + one-channel twins, never the 120 GB
expansion), so the whole table takes seconds.  The reference's TiledOrthogonalKeynet rows for AllConvNet / VGG-16 need
general-key (SpGEMM) compiles at that scale and are only produced on request (orthogonal=True; LeNet always has them)."""
import numpy as np

from . import nets, system


def count_parameters(net):
    """Parameters of a plain torch network (keynet/torch.py count_parameters)."""
    return int(sum(p.numel() for p in net.parameters() if p.requires_grad))


def parameter_table(which=('lenet', 'allconvnet'), seed=0, orthogonal=False, tiles=None, verbose=True, init=None):
    """-> list of (label, count) in the order demo/figures.py prints them.  which: any of 'lenet', 'allconvnet', 'vgg16';
    tiles: tile sizes (default: the reference's per network); init: optional callable(net) -> net setting the weights."""
    rows = []

    def add(label, n):
        rows.append((label, int(n)))
        if verbose:
            print('[figures.print_parameters]:  %s parameters=%d' % (label, int(n)))

    def keyed(factory, inshape, net, *args, **kw):
        np.random.seed(seed)
        (sensor, knet) = factory(inshape, net, *args, **kw)
        n = knet.num_parameters()
        del sensor, knet
        return n
    spec = {'lenet': ('lenet', (1, 28, 28), lambda: nets.LeNet_AvgPool(), [2, 4, 8], True),
            'allconvnet': ('allconvnet', (3, 32, 32), lambda: nets.AllConvNet(batchnorm=False), [2, 4, 8, 16], False),
            'vgg16': ('vgg-16', (3, 224, 224), lambda: nets.VGG16(), [2, 4, 8, 16, 32], False)}
    for w in which:
        (label, inshape, make, ks, small) = spec[w]
        ks = ks if tiles is None else list(tiles)
        net = make()
        net = (init(net) if init is not None else net).eval()
        add(label, count_parameters(net))
        big = {} if small else {'keep_csr': False}
        add('IdentityKeynet (%s)' % label, keyed(system.Keynet, inshape, net, **big))
        add('PermutationKeynet (%s)' % label, keyed(system.Keynet, inshape, net, global_geometric='permutation', **big))
        for k in ks:
            add('TiledPermutationKeynet-%d (%s)' % (k, label), keyed(system.TiledPermutationKeynet, inshape, net, k, **big))
        if small or orthogonal:
            for k in ([] if small and w == 'lenet' and not orthogonal else ks):
                add('TiledOrthogonalKeynet-%d (%s)' % (k, label), keyed(system.TiledOrthogonalKeynet, inshape, net, k))
    return rows


def print_parameters(which=('lenet', 'allconvnet', 'vgg16'), **kw):
    """Reference entry point name (demo/figures.py:236)."""
    return parameter_table(which=which, verbose=True, **kw)
