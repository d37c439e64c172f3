"""`dense()` factory with the reference's initialisation (backbones/dense_layer.py:63-71):
nn.Linear with fan-avg uniform weights (gain `init_scale`, 0 -> 1e-10) and zero bias."""
import math

import torch
from torch import nn


def variance_scaling_init_(tensor, scale):
    gain = 1e-10 if scale == 0 else scale
    fan_out, fan_in = tensor.shape[0], tensor.shape[1]
    rf = 1
    for d in tensor.shape[2:]:
        rf *= d
    fan_avg = (fan_in * rf + fan_out * rf) / 2.0
    bound = math.sqrt(3.0 * gain / max(1.0, fan_avg))
    with torch.no_grad():
        return tensor.uniform_(-bound, bound)


def dense(in_channels, out_channels, init_scale=1.):
    lin = nn.Linear(in_channels, out_channels)
    variance_scaling_init_(lin.weight, scale=init_scale)
    nn.init.zeros_(lin.bias)
    return lin
