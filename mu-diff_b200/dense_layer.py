"""`dense()` factory with the reference's initialisation (backbones/dense_layer.py:63-71): nn.Linear with uniform
weights (gain `init_scale`, 0 -> 1e-10) and zero bias.

The reference asks for mode='fan_avg' but its `_calculate_correct_fan` (dense_layer.py:22-32) returns
`fan_in if mode == 'fan_in' else fan_out`, i.e. 'fan_avg' silently means fan_OUT: bound = sqrt(3 * gain / fan_out).
That quirk is reproduced here so that randomly initialised modules have the reference's weight scale (it does not
matter for loaded checkpoints)."""
import math

import torch
from torch import nn


def variance_scaling_init_(tensor, scale):
    gain = 1e-10 if scale == 0 else scale
    rf = 1
    for d in tensor.shape[2:]:
        rf *= d
    fan_out = tensor.shape[0] * rf                # torch.nn.init._calculate_fan_in_and_fan_out: size(0) * receptive field
    bound = math.sqrt(3.0 * gain / max(1.0, fan_out))
    with torch.no_grad():
        return tensor.uniform_(-bound, bound)


def dense(in_channels, out_channels, init_scale=1.):
    lin = nn.Linear(in_channels, out_channels)
    variance_scaling_init_(lin.weight, scale=init_scale)
    nn.init.zeros_(lin.bias)
    return lin


def conv2d(in_planes, out_planes, kernel_size=(3, 3), stride=1, dilation=1, padding=1, bias=True, padding_mode='zeros',
           init_scale=1.):
    """backbones/dense_layer.py:73-81 (the discriminator's conv factory): an nn.Conv2d-compatible module (`weight`
    [Cout, Cin, k, k], `bias`) that runs on the libmudiff_b200 conv kernels, with the reference's initialisation."""
    from . import layers
    k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
    if stride != 1 or dilation != 1 or padding_mode != 'zeros' or padding != k // 2:
        raise NotImplementedError("mu-diff_b200 conv2d: stride 1, dilation 1, zero 'same' padding")
    conv = layers.Conv2d(in_planes, out_planes, k, bias=bias, stride=1, padding=padding)
    variance_scaling_init_(conv.weight, scale=init_scale)
    return conv
