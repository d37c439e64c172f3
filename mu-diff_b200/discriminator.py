"""Time-dependent discriminator of MU-Diff, forward pass on the libmudiff_b200 kernels:
backbones/discriminator.py:20-37 (TimestepEmbedding), :39-98 (DownConvBlock), :175-263 (Discriminator_large), with the
reference's constructor signatures, attribute names and therefore state_dict keys (the `nn.Sequential` wrappers of
:56-71 give the `conv1.0.weight` levels).

Per DownConvBlock: LeakyReLU(input) [1 elementwise kernel] -> conv3x3 + Dense(t_emb) row bias + LeakyReLU [conv epilogue]
-> FIR down-sample of both branches [fir kernel] -> conv3x3(out) + conv1x1 skip(input) as ONE K-concatenated contraction
with the 1/sqrt(2) in the epilogue.  All six `dense_t1` layers are evaluated by one GEMM.  The tail (minibatch-stddev
feature, 513-channel conv on 4x4 pixels, spatial sum, end_linear) runs in fp32.

Forward only (inference / evaluation of D); training D needs a backward pass that this package does not have.
`precision`: 'bf16' (tcgen05 tensor-core convs) or 'fp32' (CUDA-core parity path).
"""
import math

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import dense_layer, layers, ops, up_or_down_sampling

dense = dense_layer.dense
conv2d = dense_layer.conv2d
get_sinusoidal_positional_embedding = layers.get_timestep_embedding


def _slope(act):
    if not isinstance(act, nn.LeakyReLU) or abs(act.negative_slope - 0.2) > 1e-12:
        raise RuntimeError("mu-diff_b200: the discriminator kernels implement nn.LeakyReLU(0.2) (engine/train.py:474-476)")
    return 0.2


class TimestepEmbedding(nn.Module):
    def __init__(self, embedding_dim, hidden_dim, output_dim, act=nn.LeakyReLU(0.2)):
        super().__init__()
        self.embedding_dim, self.output_dim, self.hidden_dim = embedding_dim, output_dim, hidden_dim
        self.main = nn.Sequential(dense(embedding_dim, hidden_dim), act, dense(hidden_dim, output_dim))
        _slope(act)

    def forward(self, temp, act_out=L.ACT_NONE):
        temb = get_sinusoidal_positional_embedding(temp, self.embedding_dim)
        temb = ops.linear(temb, self.main[0].weight, self.main[0].bias, act_out=L.ACT_LRELU)
        return ops.linear(temb, self.main[2].weight, self.main[2].bias, act_out=act_out)


class DownConvBlock(nn.Module, layers.PackCache):
    def __init__(self, in_channel, out_channel, kernel_size=3, padding=1, t_emb_dim=128, downsample=False,
                 act=nn.LeakyReLU(0.2), fir_kernel=(1, 3, 3, 1)):
        super().__init__()
        if kernel_size != 3 or padding != 1:
            raise NotImplementedError("mu-diff_b200 DownConvBlock: 3x3 / padding 1 (the only use in the reference)")
        self.fir_kernel, self.downsample = fir_kernel, downsample
        self.conv1 = nn.Sequential(conv2d(in_channel, out_channel, kernel_size, padding=padding))
        self.conv2 = nn.Sequential(conv2d(out_channel, out_channel, kernel_size, padding=padding, init_scale=0.))
        self.dense_t1 = dense(t_emb_dim, out_channel)
        self.act = act
        self.skip = nn.Sequential(conv2d(in_channel, out_channel, 1, padding=0, bias=False))
        self.in_channel, self.out_channel = in_channel, out_channel
        _slope(act)

    def forward(self, input, t_emb, tbias=None):
        x = ops.as_nhwc(input, input.dtype if input.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        L.require_cuda(x)
        dt = x.dtype
        c1, c2, sk = self.conv1[0], self.conv2[0], self.skip[0]
        if tbias is None:
            tbias = ops.linear(t_emb, self.dense_t1.weight, self.dense_t1.bias)
        out = ops.leaky_relu(x, 0.2)
        out = ops.conv([(out, 9)], c1.packed_weight(dt), self.out_channel, bias=c1.bias_f32(), rowbias=tbias, act=L.ACT_LRELU)
        if self.downsample:
            out = up_or_down_sampling.downsample_2d(out, self.fir_kernel, factor=2)
            x = up_or_down_sampling.downsample_2d(x, self.fir_kernel, factor=2)
        wt = self._packed(('w2s', dt), [c2.weight, sk.weight],
                          lambda: torch.cat([c2.packed_weight(dt), sk.packed_weight(dt)], dim=1).contiguous())
        return ops.conv([(out, 9), (x, 1)], wt, self.out_channel, bias=c2.bias_f32(), alpha=1.0 / math.sqrt(2.0))


class Discriminator_large(nn.Module, layers.PackCache):
    """A time-dependent discriminator for large images (backbones/discriminator.py:175-263).
    forward(x, t, x_t) -> (logits [B], mid_feat [B, 8 ngf, H/8, W/8] in the compute dtype)."""

    def __init__(self, nc=1, ngf=32, t_emb_dim=128, act=nn.LeakyReLU(0.2), precision='bf16'):
        super().__init__()
        self.act = act
        _slope(act)
        self.t_embed = TimestepEmbedding(embedding_dim=t_emb_dim, hidden_dim=t_emb_dim, output_dim=t_emb_dim, act=act)
        self.start_conv = conv2d(nc, ngf * 2, 1, padding=0)
        self.conv1 = DownConvBlock(ngf * 2, ngf * 4, t_emb_dim=t_emb_dim, downsample=True, act=act)
        self.conv2 = DownConvBlock(ngf * 4, ngf * 8, t_emb_dim=t_emb_dim, downsample=True, act=act)
        self.conv3 = DownConvBlock(ngf * 8, ngf * 8, t_emb_dim=t_emb_dim, downsample=True, act=act)
        self.conv4 = DownConvBlock(ngf * 8, ngf * 8, t_emb_dim=t_emb_dim, downsample=True, act=act)
        self.conv5 = DownConvBlock(ngf * 8, ngf * 8, t_emb_dim=t_emb_dim, downsample=True, act=act)
        self.conv6 = DownConvBlock(ngf * 8, ngf * 8, t_emb_dim=t_emb_dim, downsample=True, act=act)
        self.final_conv = conv2d(ngf * 8 + 1, ngf * 8, 3, padding=1)
        self.end_linear = dense(ngf * 8, 1)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.precision = precision

    def _blocks(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6]

    def forward(self, x, t, x_t):
        L.require_cuda(x, x_t, t)
        if self.stddev_feat != 1:
            raise NotImplementedError("mu-diff_b200: stddev_feat == 1 (the reference's constant)")
        dt = torch.bfloat16 if self.precision in ('bf16', torch.bfloat16) else torch.float32
        blocks = self._blocks()
        t_embed = self.t_embed(t, act_out=L.ACT_LRELU)                                   # act(t_embed(t)), :219
        wt, bt = self._packed(('dense_t',), [p for b in blocks for p in (b.dense_t1.weight, b.dense_t1.bias)],
                              lambda: (torch.cat([b.dense_t1.weight for b in blocks], 0).float().contiguous(),
                                       torch.cat([b.dense_t1.bias for b in blocks], 0).float().contiguous()))
        tb = ops.linear(t_embed, wt, bt)                                                 # all dense_t1 rows: one GEMM
        input_x = ops.concat([x.float(), x_t.float()], dtype=torch.float32)              # torch.cat((x, x_t), 1), :221
        h = self.start_conv(input_x, compute_dtype=dt)
        off, mid_feat = 0, None
        for i, blk in enumerate(blocks):
            h = blk(h, t_embed, tbias=tb[:, off:off + blk.out_channel])
            off += blk.out_channel
            if i == 2:
                mid_feat = h                                                             # h4, :234
        batch, channel, height, width = h.shape
        group = min(batch, self.stddev_group)
        if batch % group:
            raise RuntimeError(f"shape '[{group}, -1, ...]' is invalid for a batch of {batch} (batch % group != 0)")   # :245 raises, too
        cat = ops.empty_nhwc(batch, channel + 1, height, width, torch.float32, h.device)
        ops.copy_channels(h, cat, 0)
        ops.minibatch_stddev(h, cat, channel, group)
        fc = self.final_conv
        f = ops.conv([(cat, 9)], fc.packed_weight(torch.float32), channel, bias=fc.bias_f32(), act=L.ACT_LRELU, force='simt')
        hw = height * width
        w_end = self._packed(('end', hw), [self.end_linear.weight],
                             lambda: (self.end_linear.weight.detach().float() * hw).contiguous())   # sum over HW = HW * mean
        out = ops.linear(ops.gap(f), w_end, self.end_linear.bias)
        return out.view(-1), mid_feat
