"""torch glue of the keyed path (reference: keynet/torch.py:17-113): layer-shape discovery, the
homogeneous coordinate, and conv+batchnorm fusion.  Host logic; activations may live on the GPU."""
from collections import OrderedDict

import numpy as np
import torch
from torch import nn


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def _chw(shape):
    return (int(shape[1]), int(shape[2]), int(shape[3])) if len(shape) == 4 else (int(shape[1]), 1, 1)


def netshape(net, inshape):
    """One dummy forward with hooks on every leaf child -> OrderedDict name -> {inshape, outshape,
    prevlayer, nextlayer}, bracketed by the sentinels 'input' and 'output' (keynet/torch.py:21-62).
    Shapes are canonicalised to (C,H,W); fully connected activations are (C,1,1)."""
    order = []

    def leaves(module):
        for (name, child) in module._modules.items():
            if isinstance(child, nn.Sequential):
                leaves(child)
            else:
                order.append((name, child))
    leaves(net)

    seen = []
    hooks = []

    def make_hook(name):
        def hook(m, inputs, output):
            seen.append((name, _chw(inputs[0].shape), _chw(output.shape)))
        return hook
    for (name, child) in order:
        hooks.append(child.register_forward_hook(make_hook(name)))
    net.eval()
    with torch.no_grad():
        net.forward(torch.rand(1, inshape[0], inshape[1], inshape[2]))
    for h in hooks:
        h.remove()

    d = OrderedDict()
    if len(seen) == 0:
        return d
    (first, last) = (seen[0], seen[-1])
    d['input'] = {'prevlayer': None, 'nextlayer': first[0], 'inshape': first[1], 'outshape': first[2]}
    prev = 'input'
    for (name, i, o) in seen:
        d[name] = {'inshape': i, 'outshape': o, 'prevlayer': prev, 'nextlayer': None}
        d[prev]['nextlayer'] = name
        prev = name
    d['output'] = {'nextlayer': None, 'prevlayer': last[0], 'inshape': last[1], 'outshape': last[2]}
    return d


def affine_to_linear(x):
    """N x C x H x W -> N x (C*H*W+1) with a trailing column of ones (keynet/torch.py:65-68)."""
    (N, C, H, W) = x.shape if len(x.shape) == 4 else (1, *x.shape)
    return torch.cat((x.reshape(N, C * H * W), torch.ones(N, 1, dtype=x.dtype, device=x.device)), dim=1)


def linear_to_affine(x, outshape=None):
    """N x (K+1) -> N x K, checking that the homogeneous coordinate is ~1 (atol 1e-3, raises ValueError
    otherwise; keynet/torch.py:71-77).  reshape(outshape) as in the reference."""
    assert len(x.shape) == 2
    if not bool(torch.all(torch.abs(x[:, -1].detach() - 1) <= 1E-3)):
        raise ValueError('invalid affine vector "%s"' % (str(x)))
    x_affine = torch.narrow(x, 1, 0, x.shape[1] - 1)
    return x_affine.reshape(outshape) if outshape is not None else x_affine


def affine_to_linear_matrix(W_affine, bias=None):
    """Dense (in+1) x (out+1) matrix [[W^T, 0],[b, 1]] (keynet/torch.py:80-89); the keyed path builds the
    sparse transpose directly on the GPU (sparse.keyed_linear), this dense form is kept for API parity."""
    Wt = W_affine.t()
    (R, C) = Wt.shape
    b = torch.zeros(1, C, dtype=Wt.dtype) if bias is None else bias.reshape(1, C)
    M = torch.zeros(R + 1, C + 1, dtype=Wt.dtype)
    M[:R, :C] = Wt
    M[R, :C] = b
    M[R, C] = 1
    return M


def fuse_conv2d_and_bn(conv2d_weight, conv2d_bias, bn_running_mean, bn_running_var, bn_eps, bn_weight, bn_bias):
    """Fold an eval-mode BatchNorm2d into the preceding conv (keynet/torch.py:99-113):
    w' = w * g/sqrt(var+eps),  b' = (b-mean)/sqrt(var+eps)*g + beta, evaluated in this order in fp32."""
    std = torch.sqrt(bn_running_var + np.float32(bn_eps))
    b = conv2d_bias if conv2d_bias is not None else bn_running_mean.new_zeros(bn_running_mean.shape)
    w = conv2d_weight * (bn_weight / std).reshape([conv2d_weight.shape[0], 1, 1, 1])
    b = (((b - bn_running_mean) / std) * bn_weight) + bn_bias
    return (w, b)
