"""Plain (un-keyed) torch networks that are the INPUTS of the keyed path: the architectures the reference
ships in keynet/mnist.py:11-63, keynet/cifar10.py:12-65 and keynet/vgg.py:38-122.  Layer names matter: the
key chaining keys layers by name and merges 'xyz_bn' / '*relu*' layers into their predecessor."""
from collections import OrderedDict

import torch
from torch import nn


def _sequential_forward(net, x, flatten_before):
    for (name, m) in net.named_children():
        if name in flatten_before:
            x = x.reshape(x.shape[0], -1)
        x = m(x)
    return x


class LeNet(nn.Module):
    """LeNet with padding, odd filters and even image sizes; max pooling (not keyable, baseline only)."""
    _pool = staticmethod(lambda: nn.MaxPool2d(kernel_size=3, stride=2, padding=1))

    def __init__(self):
        super(LeNet, self).__init__()
        spec = OrderedDict([
            ('conv1', nn.Conv2d(1, 6, kernel_size=3, stride=1, padding=1)), ('relu1', nn.ReLU()), ('pool1', self._pool()),
            ('conv2', nn.Conv2d(6, 16, kernel_size=3, stride=1, padding=1)), ('relu2', nn.ReLU()), ('pool2', self._pool()),
            ('fc1', nn.Linear(7 * 7 * 16, 120)), ('relu3', nn.ReLU()),
            ('fc2', nn.Linear(120, 84)), ('relu4', nn.ReLU()),
            ('fc3', nn.Linear(84, 10))])
        for (k, m) in spec.items():
            setattr(self, k, m)

    def forward(self, x):
        return _sequential_forward(self, x, ('fc1',))


class LeNet_AvgPool(LeNet):
    """1x28x28 -> 10; conv(6)-relu-avgpool(3,2,1)-conv(16)-relu-avgpool(3,2,1)-fc120-relu-fc84-relu-fc10."""
    _pool = staticmethod(lambda: nn.AvgPool2d(kernel_size=3, stride=2, padding=1))


class AllConvNet(nn.Module):
    """All-convolutional CIFAR-10 net, 3x32x32 -> 10 (dropout layers are the identity in eval mode)."""

    def __init__(self, batchnorm=False, n_input_channels=3, n_classes=10, **kwargs):
        super(AllConvNet, self).__init__()
        self._batchnorm = batchnorm
        L = OrderedDict()
        L['dropout0'] = nn.Dropout(p=0.2)
        L['conv1'] = nn.Conv2d(n_input_channels, 96, 3, padding=1); L['relu1'] = nn.ReLU()
        L['conv2'] = nn.Conv2d(96, 96, 3, padding=1); L['relu2'] = nn.ReLU()
        L['conv3'] = nn.Conv2d(96, 96, 3, padding=1, stride=2)
        if batchnorm:
            L['conv3_bn'] = nn.BatchNorm2d(96)
        L['dropout3'] = nn.Dropout(p=0.5); L['relu3'] = nn.ReLU()
        L['conv4'] = nn.Conv2d(96, 192, 3, padding=1); L['relu4'] = nn.ReLU()
        L['conv5'] = nn.Conv2d(192, 192, 3, padding=1); L['relu5'] = nn.ReLU()
        L['conv6'] = nn.Conv2d(192, 192, 3, padding=1, stride=2)
        if batchnorm:
            L['conv6_bn'] = nn.BatchNorm2d(192)
        L['dropout6'] = nn.Dropout(p=0.5); L['relu6'] = nn.ReLU()
        L['conv7'] = nn.Conv2d(192, 192, 3, padding=1); L['relu7'] = nn.ReLU()
        L['conv8'] = nn.Conv2d(192, 192, 1); L['relu8'] = nn.ReLU()
        L['conv9'] = nn.Conv2d(192, n_classes, 1); L['relu9'] = nn.ReLU()
        L['fc1'] = nn.Linear(n_classes * 8 * 8, 100); L['relu10'] = nn.ReLU()
        L['fc2'] = nn.Linear(100, 10)
        for (k, m) in L.items():
            setattr(self, k, m)

    def forward(self, x):
        return _sequential_forward(self, x, ('fc1',))


class VGG16(nn.Module):
    """VGG-16 (3x224x224) with average pooling, as keyed by the reference (keynet/vgg.py:38-122)."""

    def __init__(self, num_classes=2622, avgpool=True):
        super(VGG16, self).__init__()
        pool = (lambda: nn.AvgPool2d((3, 3), (2, 2), (0, 0), ceil_mode=True)) if avgpool else (lambda: nn.MaxPool2d((2, 2), (2, 2), (0, 0), ceil_mode=True))
        cfg = [(1, [(3, 64), (64, 64)]), (2, [(64, 128), (128, 128)]), (3, [(128, 256), (256, 256), (256, 256)]),
               (4, [(256, 512), (512, 512), (512, 512)]), (5, [(512, 512), (512, 512), (512, 512)])]
        for (b, convs) in cfg:
            for (i, (cin, cout)) in enumerate(convs, start=1):
                setattr(self, 'conv%d_%d' % (b, i), nn.Conv2d(cin, cout, (3, 3), (1, 1), (1, 1)))
                setattr(self, 'relu%d_%d' % (b, i), nn.ReLU())
            setattr(self, 'pool%d_%d' % (b, len(convs)), pool())
        self.fc6 = nn.Linear(25088, 4096); self.relu6 = nn.ReLU()
        self.dropout7 = nn.Dropout(0.5); self.fc7 = nn.Linear(4096, 4096); self.relu7 = nn.ReLU()
        self.dropout8 = nn.Dropout(0.5); self.fc8 = nn.Linear(4096, num_classes)

    def forward(self, x):
        assert x.ndim == 4 and tuple(x.shape[1:]) == (3, 224, 224), "Invalid input shape - must be Nx3x224x224"
        return _sequential_forward(self, x, ('fc6',))
