"""CPU oracle for the EdgeConv layer -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).

Restates, in float64 on the CPU, the reference's op sequence
    get_graph_features   src/utils/neighbour_ops.py:113-119   cat(x_j - x_i, x_i) over the k neighbours -> (B,2C,N,k)
    EdgeConvLayer        src/module/layers.py:159-203         Conv2d 1x1 (no bias) -> BatchNorm2d -> activation
    max over k           src/module/encoders.py:52
literally (the edge tensor IS materialised here), with BatchNorm2d written out: batch statistics over (B,N,k) with the
biased variance for normalisation and the unbiased one for the running estimate (momentum update).
Pinned against tests/golden/edgeconv.npz, which tests/golden/make_golden.py produced by running the reference's own
layers.py / neighbour_ops.py.
"""
from __future__ import annotations

import torch


def edge_conv_max(x, idx, weight, gamma, beta, running_mean=None, running_var=None, training=True, momentum=0.1,
                  eps=1e-5, negative_slope=None):
    """x (B,C,N), idx (B,N,k) int64, weight (Cout,2C) -> out (B,Cout,N) float64 (autograd-capable), and the updated
    running statistics (or None)."""
    x = x.double()
    b, c, n = x.shape
    k = idx.shape[2]
    flat = idx.reshape(b, 1, n * k).expand(-1, c, -1)
    nbr = torch.gather(x, 2, flat).view(b, c, n, k)                     # neighbour_ops.py:92-93
    ctr = x.unsqueeze(3).expand(-1, -1, -1, k)
    feat = torch.cat([nbr - ctr, ctr], dim=1)                           # neighbour_ops.py:116-117
    y = torch.einsum("oc,bcnk->bonk", weight.double(), feat)            # Conv2d 1x1, bias=False (layers.py:199)
    new_rm = new_rv = None
    if training:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if running_mean is not None:
            e = b * n * k
            new_rm = (1 - momentum) * running_mean.double() + momentum * mean.detach()
            new_rv = (1 - momentum) * running_var.double() + momentum * var.detach() * e / max(e - 1, 1)
    else:
        mean, var = running_mean.double(), running_var.double()
    z = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    z = z * gamma.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
    if negative_slope is not None:
        z = torch.where(z > 0, z, z * negative_slope)                   # LeakyReLU / ReLU (slope 0)
    return z.max(dim=3)[0], new_rm, new_rv
