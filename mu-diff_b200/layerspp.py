"""NCSN++ building blocks of the MU-Diff generators with the reference's constructor /
forward signatures and state_dict keys (backbones/layerspp.py), executed by the
libmudiff_b200 kernels:

  AdaptiveGroupNorm (:37-54)  GroupNorm_Conv (:56-65)  Combine (:80-95)  AttnBlockpp (:98-137)
  Upsample (:141-173)  Downsample (:176-210)  ResnetBlockBigGANpp_Adagn (:261-324)
  ConvFeatBlock (:394-423)  ConvBlock (:426-455)  ConvBlock_GAP (:458-501)

Tensors are logical NCHW with channels-last memory; the dtype of the incoming activation
selects the path: bfloat16 -> tcgen05 tensor-core kernels, float32 -> CUDA-core fp32 kernels.
A block input may be a tuple `(h, skip)`: the channel concat torch.cat([h, skip], 1) of
ncsnpp_generator_adagn_feat.py:383 is then never materialised (GroupNorm reads both
sources, the fused 1x1 shortcut takes them as separate K segments).
"""
import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import dense_layer, layers, ops, up_or_down_sampling

conv1x1 = layers.ddpm_conv1x1
conv3x3 = layers.ddpm_conv3x3
NIN = layers.NIN
default_init = layers.default_init
dense = dense_layer.dense

_F = (torch.float32, torch.bfloat16)


def _srcs(x):
    if isinstance(x, (tuple, list)):
        return [ops.as_nhwc(t) for t in x]
    return [ops.as_nhwc(x)]


def _compute_dtype(t):
    return t.dtype if t.dtype in _F else torch.float32


def _stem_fusable(conv, x):
    """1 -> N stem conv whose GroupNorm + activation can be folded into the conv kernel (ops.stem_conv_gn_act)."""
    return (ops.FUSED_STEM and conv.in_channels == 1 and x.shape[1] == 1 and conv.kernel_size == 3
            and conv.out_channels % 8 == 0 and conv.out_channels <= 256)


class AdaptiveGroupNorm(nn.Module):
    def __init__(self, num_groups, in_channel, style_dim):
        super().__init__()
        self.norm = nn.GroupNorm(num_groups, in_channel, affine=False, eps=1e-6)
        self.style = dense(style_dim, in_channel * 2)
        self.style.bias.data[:in_channel] = 1
        self.style.bias.data[in_channel:] = 0
        self.num_groups, self.in_channel = num_groups, in_channel

    def style_params(self, style):
        """Linear(style) -> [B, 2C] fp32: gamma = [:, :C], beta = [:, C:]."""
        return ops.linear(style, self.style.weight, self.style.bias)

    def scale_shift(self, input, style, gb=None):
        """(scale, shift) table [B, C, 2] of this AdaGN for consumers that apply it themselves (ops.conv xform)."""
        srcs = _srcs(input)
        L.require_cuda(*srcs)
        gb = self.style_params(style) if gb is None else gb
        c = self.in_channel
        return ops.gn_scale_shift(srcs, None, self.num_groups, gamma=gb, beta=gb[:, c:],
                                  gb_bstride=gb.stride(0), eps=self.norm.eps)

    def forward(self, input, style, act=L.ACT_NONE, gb=None):
        srcs = _srcs(input)
        L.require_cuda(*srcs)
        gb = self.style_params(style) if gb is None else gb
        c = self.in_channel
        return ops.gn_apply_auto(srcs, self.num_groups, gamma=gb, beta=gb[:, c:], gb_bstride=gb.stride(0),
                                 eps=self.norm.eps, act=act)


class GroupNorm_Conv(nn.Module):
    def __init__(self, num_groups, in_channel):
        super().__init__()
        self.norm = nn.GroupNorm(num_groups, in_channel, affine=False, eps=1e-6)
        self.num_groups = num_groups

    def forward(self, input, act=L.ACT_NONE):
        return ops.group_norm(_srcs(input), self.num_groups, eps=self.norm.eps, act=act)


class GroupNorm(nn.GroupNorm):
    """nn.GroupNorm (affine) with the CUDA path of this package; keeps `weight`/`bias` keys."""

    def forward(self, input, act=L.ACT_NONE):
        return ops.group_norm(_srcs(input), self.num_groups, gamma=self.weight, beta=self.bias, gb_bstride=0,
                              eps=self.eps, act=act)


class GaussianFourierProjection(nn.Module):
    """Constructor kept importable (layerspp.py:68-77); not reachable with embedding_type='positional'."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size) * scale, requires_grad=False)

    def forward(self, x):
        x_proj = x[:, None] * self.W[None, :] * 2 * np.pi
        return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


class Combine(nn.Module):
    """Combine information from skip connections (layerspp.py:80-95)."""

    def __init__(self, dim1, dim2, method='cat'):
        super().__init__()
        self.Conv_0 = conv1x1(dim1, dim2)
        self.method = method

    def forward(self, x, y):
        y = ops.as_nhwc(y)
        if self.method == 'cat':
            b, c2, h, w = y.shape
            dim2 = self.Conv_0.out_channels
            out = ops.empty_nhwc(b, dim2 + c2, h, w, y.dtype, y.device)
            self.Conv_0(x, compute_dtype=y.dtype, out=out, out_coff=0)
            ops.copy_channels(y, out, dim2)
            return out
        elif self.method == 'sum':
            return self.Conv_0(x, compute_dtype=y.dtype, residual=y, alpha=1.0, beta=1.0)
        else:
            raise ValueError(f'Method {self.method} not recognized.')


class AttnBlockpp(nn.Module, layers.PackCache):
    """Channel-wise self-attention block (layerspp.py:98-137).

    GN(affine) -> [q|k] = one GEMM (N = 2C) ; V^T = swapped GEMM (weights as the M operand) ;
    S = q k^T * C^-1/2 ; row softmax ; O = P V ; out = (x + O W3 + b3') * 1/sqrt(2)
    with the V bias folded into b3' = b3 + b2 W3 (softmax rows sum to one)."""

    def __init__(self, channels, skip_rescale=False, init_scale=0.):
        super().__init__()
        self.GroupNorm_0 = GroupNorm(num_groups=min(channels // 4, 32), num_channels=channels, eps=1e-6)
        self.NIN_0 = NIN(channels, channels)
        self.NIN_1 = NIN(channels, channels)
        self.NIN_2 = NIN(channels, channels)
        self.NIN_3 = NIN(channels, channels, init_scale=init_scale)
        self.skip_rescale = skip_rescale

    def forward(self, x):
        x = ops.as_nhwc(x, _compute_dtype(x))
        L.require_cuda(x)
        B, C, H, W = x.shape
        Lt = H * W
        dt = x.dtype
        n0, n1, n2, n3 = self.NIN_0, self.NIN_1, self.NIN_2, self.NIN_3
        w_qk = self._packed(('qk', dt), [n0.W, n1.W], lambda: torch.cat([n0.W.t(), n1.W.t()], 0).to(dt).contiguous())
        b_qk = self._packed(('bqk',), [n0.b, n1.b], lambda: torch.cat([n0.b, n1.b]).float().contiguous())
        w_v = self._packed(('v', dt), [n2.W], lambda: n2.W.detach().t().to(dt).contiguous().view(1, 1, C, C).permute(0, 3, 1, 2))
        w_o = n3.packed_weight(dt)
        b_o = self._packed(('bo',), [n2.b, n3.b, n3.W], lambda: (n3.b + n2.b @ n3.W).float().contiguous())

        hn = self.GroupNorm_0(x)                                           # [B,C,H,W]
        hn_flat = hn.permute(0, 2, 3, 1).reshape(B, 1, Lt, C).permute(0, 3, 1, 2)   # logical [B,C,1,L]
        qk = ops.conv([(hn_flat, 1)], w_qk, 2 * C, bias=b_qk, pad=0)      # [B,2C,1,L]
        # V^T[b] = W2^T (M = C rows) x hn[b]^T : A = weights (shared), "weights" = hn (per sample)
        vt = ops.conv([(w_v, 1)], hn, Lt, pad=0, a_batched=False, batch=B, w_bstride=Lt * C, w_ld=C)   # [B,L,1,C] == V^T [B][C][L]
        if ops.attention_supported(C, Lt, dt):
            # fused QK^T -> softmax -> PV (flash-style, tcgen05): the [L, L] scores never reach HBM
            o = ops.attention(qk, vt, B, Lt, C, float(int(C) ** (-0.5)))   # [B,L,C]
            o = o.view(B, H, W, C).permute(0, 3, 1, 2)
        else:
            q = qk[:, :C]
            k_ptr_view = qk[:, C:]
            s = ops.conv([(q, 1)], k_ptr_view, Lt, pad=0, alpha=float(int(C) ** (-0.5)),
                         w_bstride=Lt * 2 * C, w_ld=2 * C)                     # [B,L,1,L] scores
            ops.softmax_rows_(s.permute(0, 2, 3, 1).reshape(B * Lt, Lt))
            o = ops.conv([(s, 1)], vt, C, pad=0, w_bstride=C * Lt, w_ld=Lt)   # [B,C,1,L]
            o = o.permute(0, 2, 3, 1).reshape(B, H, W, C).permute(0, 3, 1, 2)
        sc = ops.SQRT2_INV if self.skip_rescale else 1.0
        return ops.conv([(o, 1)], w_o, C, bias=b_o, pad=0, residual=x, alpha=sc, beta=sc, want_stats=True)


class Upsample(nn.Module):
    def __init__(self, in_ch=None, out_ch=None, with_conv=False, fir=False, fir_kernel=(1, 3, 3, 1)):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        if not fir:
            if with_conv:
                self.Conv_0 = conv3x3(in_ch, out_ch)
        else:
            if with_conv:
                self.Conv2d_0 = up_or_down_sampling.Conv2d(in_ch, out_ch, kernel=3, up=True,
                                                           resample_kernel=fir_kernel, use_bias=True,
                                                           kernel_init=default_init())
        self.fir, self.with_conv, self.fir_kernel, self.out_ch = fir, with_conv, fir_kernel, out_ch

    def forward(self, x):
        if not self.fir:
            h = up_or_down_sampling.naive_upsample_2d(x, 2)       # == F.interpolate(nearest, x2)
            if self.with_conv:
                h = self.Conv_0(h)
        else:
            if not self.with_conv:
                h = up_or_down_sampling.upsample_2d(x, self.fir_kernel, factor=2)
            else:
                h = self.Conv2d_0(x)
        return h


class Downsample(nn.Module):
    def __init__(self, in_ch=None, out_ch=None, with_conv=False, fir=False, fir_kernel=(1, 3, 3, 1)):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        if not fir:
            if with_conv:
                self.Conv_0 = conv3x3(in_ch, out_ch, stride=2, padding=0)
        else:
            if with_conv:
                self.Conv2d_0 = up_or_down_sampling.Conv2d(in_ch, out_ch, kernel=3, down=True,
                                                           resample_kernel=fir_kernel, use_bias=True,
                                                           kernel_init=default_init())
        self.fir, self.fir_kernel, self.with_conv, self.out_ch = fir, fir_kernel, with_conv, out_ch

    def forward(self, x, compute_dtype=None):
        if not self.fir:
            if self.with_conv:
                raise NotImplementedError("non-FIR strided conv down-sampling is not on the MU-Diff sampling path")
            return up_or_down_sampling.naive_downsample_2d(x, 2)  # == avg_pool2d(2)
        if not self.with_conv:
            return up_or_down_sampling.downsample_2d(x, self.fir_kernel, factor=2)
        return self.Conv2d_0(x, compute_dtype=compute_dtype)


class ResnetBlockBigGANpp_Adagn(nn.Module, layers.PackCache):
    def __init__(self, act, in_ch, out_ch=None, temb_dim=None, zemb_dim=None, up=False, down=False,
                 dropout=0.1, fir=False, fir_kernel=(1, 3, 3, 1), skip_rescale=True, init_scale=0.):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        self.GroupNorm_0 = AdaptiveGroupNorm(min(in_ch // 4, 32), in_ch, zemb_dim)
        self.up, self.down, self.fir, self.fir_kernel = up, down, fir, fir_kernel
        self.Conv_0 = conv3x3(in_ch, out_ch)
        if temb_dim is not None:
            self.Dense_0 = nn.Linear(temb_dim, out_ch)
            self.Dense_0.weight.data = default_init()(self.Dense_0.weight.shape)
            nn.init.zeros_(self.Dense_0.bias)
        self.GroupNorm_1 = AdaptiveGroupNorm(min(out_ch // 4, 32), out_ch, zemb_dim)
        self.Dropout_0 = nn.Dropout(dropout)
        self.Conv_1 = conv3x3(out_ch, out_ch, init_scale=init_scale)
        if in_ch != out_ch or up or down:
            self.Conv_2 = conv1x1(in_ch, out_ch)
        self.skip_rescale = skip_rescale
        self.act = act
        self.in_ch, self.out_ch = in_ch, out_ch

    def _resample(self, t):
        if self.up:
            return (up_or_down_sampling.upsample_2d(t, self.fir_kernel, factor=2) if self.fir
                    else up_or_down_sampling.naive_upsample_2d(t, factor=2))
        return (up_or_down_sampling.downsample_2d(t, self.fir_kernel, factor=2) if self.fir
                else up_or_down_sampling.naive_downsample_2d(t, factor=2))

    def forward(self, x, temb=None, zemb=None, gb0=None, gb1=None, tbias=None):
        """x: tensor or (h, skip) tuple.  gb0/gb1/tbias: optional pre-computed AdaGN style rows /
        Dense_0(act(temb)) (the generators batch all of them into one GEMM per step)."""
        if not isinstance(self.act, nn.SiLU):
            raise RuntimeError("mu-diff_b200: only nn.SiLU activations are fused (the generators use nn.SiLU)")
        if self.training and self.Dropout_0.p > 0:
            raise RuntimeError("mu-diff_b200: inference path only (dropout is active)")
        xs = _srcs(x)
        L.require_cuda(*xs)
        dt = _compute_dtype(xs[0])
        xs = [t if t.dtype == dt else t.to(dt) for t in xs]
        seg_c = [t.shape[1] for t in xs]
        assert sum(seg_c) == self.in_ch, (seg_c, self.in_ch)

        # bf16 tensor-core path: AdaGN + SiLU can be applied by the conv kernel to its staged operand tiles
        # (ops.conv xform) - the normalised tensor is then never stored.  Not across a FIR resample (the FIR reads
        # it), and only where it pays (ops.xform_profitable).
        tc_ok = dt == torch.bfloat16 and all(c % 64 == 0 for c in seg_c) and self.out_ch % 64 == 0
        b_, _, h_, w_ = xs[0].shape
        pix0 = b_ * h_ * w_
        pix1 = pix0 * 4 if self.up else (pix0 // 4 if self.down else pix0)       # Conv_1 runs at the resampled resolution
        fuse0 = tc_ok and not (self.up or self.down) and ops.xform_profitable(seg_c, pixels=pix0)
        fuse1 = tc_ok and ops.xform_profitable([self.out_ch], len(xs) if hasattr(self, 'Conv_2') else 0, pixels=pix1)
        if tbias is None and temb is not None:
            tbias = ops.linear(temb, self.Dense_0.weight, self.Dense_0.bias, act_in=L.ACT_SILU)
        if fuse0:
            tab0 = self.GroupNorm_0.scale_shift(tuple(xs), zemb, gb=gb0)
            offs = [0] + [sum(seg_c[:i + 1]) for i in range(len(seg_c) - 1)]
            h = ops.conv([(t, 9, (tab0, o)) for t, o in zip(xs, offs)], self.Conv_0.packed_weight(dt, seg_c), self.out_ch,
                         bias=self.Conv_0.bias_f32(), rowbias=tbias, want_stats=True)
        else:
            fused = None
            if (self.up or self.down) and self.fir and len(xs) == 1:
                # AdaGN + SiLU + FIR of h AND the FIR of x from one read of x (layerspp.py:293-305): the full-resolution
                # normalised tensor is never written
                tab0 = self.GroupNorm_0.scale_shift(xs[0], zemb, gb=gb0)
                fused = up_or_down_sampling.resample_2d_gn(xs[0], tab0, self.fir_kernel, up=self.up)
            if fused is not None:
                h, xr = fused
                xs = [xr]
            else:
                h = self.GroupNorm_0(tuple(xs), zemb, act=L.ACT_SILU, gb=gb0)      # AdaGN + SiLU, one pass
                if self.up or self.down:
                    h = self._resample(h)
                    xs = [self._resample(t) for t in xs]
            h = ops.conv([(h, 9)], self.Conv_0.packed_weight(dt), self.out_ch, bias=self.Conv_0.bias_f32(),
                         rowbias=tbias, want_stats=True)
        if fuse1:
            hseg = (h, 9, (self.GroupNorm_1.scale_shift(h, zemb, gb=gb1), 0))
        else:
            hseg = (self.GroupNorm_1(h, zemb, act=L.ACT_SILU, gb=gb1), 9)
        sc = ops.SQRT2_INV if self.skip_rescale else 1.0
        w1 = self.Conv_1.packed_weight(dt)
        if hasattr(self, 'Conv_2') and dt == torch.float32 and ops.FP32_TC and self.in_ch % 64 == 0 and self.out_ch % 64 == 0:
            # fp32 path on the tensor cores (ops.conv splits a SINGLE fp32 segment into three bf16 ones): the shortcut is
            # its own contraction and enters Conv_1's epilogue as the residual
            c2 = self.Conv_2
            xcat = xs[0] if len(xs) == 1 else ops.concat(xs)
            short = ops.conv([(xcat, 1)], c2.packed_weight(dt), self.out_ch, bias=c2.bias_f32(), pad=0)
            return ops.conv([hseg], w1, self.out_ch, bias=self.Conv_1.bias_f32(), residual=short, alpha=sc, beta=sc,
                            want_stats=True)
        if hasattr(self, 'Conv_2'):
            # Conv_1(h) + Conv_2(x) as ONE contraction: K = 9*Cout + Cin   (layerspp.py:316-324)
            c2 = self.Conv_2
            wt = self._packed(('w12', dt, tuple(seg_c)), [self.Conv_1.weight, c2.weight],
                              lambda: torch.cat([self.Conv_1.packed_weight(dt), c2.packed_weight(dt, seg_c)], dim=1).contiguous())
            bias = self._packed(('b12',), [self.Conv_1.bias, c2.bias],
                                lambda: (self.Conv_1.bias + c2.bias).detach().float().contiguous())
            return ops.conv([hseg] + [(t, 1) for t in xs], wt, self.out_ch, bias=bias, alpha=sc, want_stats=True)
        res = xs[0] if len(xs) == 1 else ops.concat(xs)
        return ops.conv([hseg], w1, self.out_ch, bias=self.Conv_1.bias_f32(), residual=res, alpha=sc, beta=sc,
                        want_stats=True)


class ConvFeatBlock(nn.Module):
    """conv3x3 -> GroupNorm(affine=False) -> act -> conv3x3 (layerspp.py:394-423)."""

    def __init__(self, act, in_ch=None, out_ch=None, zemb_dim=256):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.group_norm = GroupNorm_Conv(min(out_ch // 4, 32), out_ch)
        self.act = act
        self.conv2 = conv3x3(out_ch, out_ch)

    def forward(self, x, compute_dtype=None, out=None, out_coff=0, stats_out=None):
        dt = compute_dtype or _compute_dtype(x)
        if _stem_fusable(self.conv1, x):
            h = ops.stem_conv_gn_act(x, self.conv1.packed_weight(torch.float32), self.conv1.bias_f32(),
                                     self.group_norm.num_groups, eps=self.group_norm.norm.eps, act=L.ACT_SILU, out_dtype=dt)
        else:
            h = self.conv1(x, compute_dtype=dt, want_stats=True)
            h = self.group_norm(h, act=L.ACT_SILU)
        return self.conv2(h, compute_dtype=dt, out=out, out_coff=out_coff, want_stats=True, stats_out=stats_out)


class ConvBlock(nn.Module):
    """conv3x3 -> AdaGN(style) -> act -> conv3x3 (layerspp.py:426-455)."""

    def __init__(self, act, in_ch=None, out_ch=None, zemb_dim=256):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.group_norm = AdaptiveGroupNorm(min(out_ch // 4, 32), out_ch, zemb_dim)
        self.act = act
        self.conv2 = conv3x3(out_ch, out_ch)

    def forward(self, x, style=None, compute_dtype=None, out=None, out_coff=0):
        dt = compute_dtype or _compute_dtype(x)
        if _stem_fusable(self.conv1, x):
            gn = self.group_norm
            gb = gn.style_params(style)
            h = ops.stem_conv_gn_act(x, self.conv1.packed_weight(torch.float32), self.conv1.bias_f32(), gn.num_groups,
                                     gamma=gb, beta=gb[:, gn.in_channel:], gb_bstride=gb.stride(0), eps=gn.norm.eps,
                                     act=L.ACT_SILU, out_dtype=dt)
        else:
            h = self.conv1(x, compute_dtype=dt, want_stats=True)
            h = self.group_norm(h, style, act=L.ACT_SILU)
        return self.conv2(h, compute_dtype=dt, out=out, out_coff=out_coff)


class ConvBlock_GAP(nn.Module):
    """conv3x3 -> GN -> act -> conv3x3 -> global average pool -> dense (layerspp.py:458-501)."""

    def __init__(self, act, in_ch=None, out_ch=None, zemb_dim=256):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.group_norm = GroupNorm_Conv(min(out_ch // 4, 32), out_ch)
        self.act = act
        self.conv2 = conv3x3(out_ch, out_ch)
        self.adaptive_gap = nn.AdaptiveAvgPool2d(1)
        self.fc = dense(out_ch, zemb_dim)

    def forward(self, x, compute_dtype=None):
        dt = compute_dtype or _compute_dtype(x)
        if _stem_fusable(self.conv1, x):
            h = ops.stem_conv_gn_act(x, self.conv1.packed_weight(torch.float32), self.conv1.bias_f32(),
                                     self.group_norm.num_groups, eps=self.group_norm.norm.eps, act=L.ACT_SILU, out_dtype=dt)
        else:
            h = self.conv1(x, compute_dtype=dt, want_stats=True)
            h = self.group_norm(h, act=L.ACT_SILU)
        h = self.conv2(h, compute_dtype=dt)
        g = ops.gap(h)
        assert g.shape[1] == self.fc.in_features, f"GAP vector {g.shape[1]} != fc.in_features {self.fc.in_features}"
        return ops.linear(g, self.fc.weight, self.fc.bias)
