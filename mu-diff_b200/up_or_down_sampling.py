"""FIR resampling front-end (backbones/up_or_down_sampling.py): `_setup_kernel` (:186-193),
`upsample_2d` (:200-229), `downsample_2d` (:232-262), `conv_downsample_2d` (:149-183),
StyleGAN2 `Conv2d` with fused down-sampling (:28-61), `naive_*` (:64-74).
Pads / gains are computed exactly as the reference; the FIR itself is `mudiff_upfirdn2d`.
The device copy of the tiny kernel is cached (no per-call H2D -> CUDA-graph safe)."""
import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from .layers import PackCache
from .op import upfirdn2d


def _setup_kernel(k):
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    assert k.ndim == 2
    assert k.shape[0] == k.shape[1]
    return k


def _fir(x, k_np, up=1, down=1, pad=(0, 0)):
    """channels-last fast path when possible, reference-compatible `upfirdn2d` otherwise."""
    kdev = ops.fir_kernel_device(np.ascontiguousarray(k_np, dtype=np.float32), x.device)
    if x.requires_grad and torch.is_grad_enabled():
        return upfirdn2d(x, kdev.to(x.dtype), up=up, down=down, pad=pad)
    return ops.upfirdn2d_nhwc(x, kdev, up=up, down=down, pad=pad)


def upsample_2d(x, k=None, factor=2, gain=1):
    assert isinstance(factor, int) and factor >= 1
    if k is None:
        k = [1] * factor
    k = _setup_kernel(k) * (gain * (factor ** 2))
    p = k.shape[0] - factor
    return _fir(x, k, up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=None, factor=2, gain=1):
    assert isinstance(factor, int) and factor >= 1
    if k is None:
        k = [1] * factor
    k = _setup_kernel(k) * gain
    p = k.shape[0] - factor
    return _fir(x, k, down=factor, pad=((p + 1) // 2, p // 2))


def resample_2d_gn(x, table, k=None, up=False, factor=2, gain=1):
    """(upsample_2d | downsample_2d)(act(AdaGN(x))) AND the same resampling of x itself, fused (ops.fir_resample_gn);
    kernel / gain / pads exactly as upsample_2d (:200-229) / downsample_2d (:232-262).  None if not applicable."""
    if factor != 2 or (x.requires_grad and torch.is_grad_enabled()):
        return None
    kk = _setup_kernel([1] * factor if k is None else k) * (gain * (factor ** 2) if up else gain)
    if kk.shape != (4, 4):
        return None
    kdev = ops.fir_kernel_device(np.ascontiguousarray(kk, dtype=np.float32), x.device)
    return ops.fir_resample_gn(x, table, kdev, up)


def naive_upsample_2d(x, factor=2):
    _N, C, H, W = x.shape
    k = np.ones((factor, factor), dtype=np.float32)          # nearest neighbour == zero-insert * ones
    return _fir(x, k, up=factor, pad=(factor - 1, 0))


def naive_downsample_2d(x, factor=2):
    k = np.ones((factor, factor), dtype=np.float32) / (factor * factor)
    return _fir(x, k, down=factor, pad=(0, 0))


def conv_downsample_2d(x, w, k=None, factor=2, gain=1, _packed=None, _bias=None, _out_dtype=None):
    """FIR pre-filter (pad ((p+1)//2, p//2)) then stride-`factor` VALID conv (:149-183)."""
    assert isinstance(factor, int) and factor >= 1
    _outC, _inC, convH, convW = w.shape
    assert convW == convH
    if k is None:
        k = [1] * factor
    k = _setup_kernel(k) * gain
    p = (k.shape[0] - factor) + (convW - 1)
    x = _fir(x, k, pad=((p + 1) // 2, p // 2))
    wt = _packed if _packed is not None else ops.pack_conv_weight(w, (_inC,), x.dtype)
    if (factor == 2 and convH == 3 and x.dtype == torch.bfloat16 and _inC % 64 == 0 and _outC % 32 == 0
            and x.shape[2] % 2 == 1 and x.shape[3] % 2 == 1):
        # tensor-core path: 'same' conv, keep the odd outputs (4x the FLOPs of the strided conv, still ~8x faster
        # than the CUDA-core kernel: K = 576 is far too small to matter next to the 3x3 convs of the level)
        return ops.conv([(x, 9)], wt, _outC, bias=_bias, pad=1, dec2=True, out_dtype=_out_dtype)
    return ops.conv([(x, convH * convW)], wt, _outC, bias=_bias, stride=factor, pad=0, force='simt',
                    out_dtype=_out_dtype)


def upsample_conv_2d(x, w, k=None, factor=2, gain=1):
    # The reference implementation is unreachable and broken (w[..., ::-1, ::-1] on a torch tensor,
    # backbones/up_or_down_sampling.py:131); it is not on the sampling path.
    raise NotImplementedError("upsample_conv_2d is not on the MU-Diff sampling path (broken in the reference)")


class Conv2d(nn.Module, PackCache):
    """StyleGAN2-style conv with optional fused down-sampling (:28-61); used by
    layerspp.Downsample(with_conv=True, fir=True) for the input pyramid."""

    def __init__(self, in_ch, out_ch, kernel, up=False, down=False, resample_kernel=(1, 3, 3, 1), use_bias=True,
                 kernel_init=None):
        super().__init__()
        assert not (up and down)
        assert kernel >= 1 and kernel % 2 == 1
        self.weight = nn.Parameter(torch.zeros(out_ch, in_ch, kernel, kernel))
        if kernel_init is not None:
            self.weight.data = kernel_init(self.weight.data.shape)
        if use_bias:
            self.bias = nn.Parameter(torch.zeros(out_ch))
        self.up, self.down = up, down
        self.resample_kernel = resample_kernel
        self.kernel = kernel
        self.use_bias = use_bias

    def forward(self, x, compute_dtype=None):
        L.require_cuda(x)
        dt = compute_dtype or (x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        in_ch = self.weight.shape[1]
        wdt = torch.float32 if in_ch < 8 else dt
        x = ops.as_nhwc(x, wdt)
        wt = self._packed(('w', wdt), [self.weight], lambda: ops.pack_conv_weight(self.weight, (in_ch,), wdt))
        bias = self._packed(('b',), [self.bias], lambda: self.bias.detach().float().contiguous()) if self.use_bias else None
        if self.up:
            return upsample_conv_2d(x, self.weight, k=self.resample_kernel)
        if self.down:
            y = conv_downsample_2d(x, self.weight, k=self.resample_kernel, _packed=wt, _bias=bias, _out_dtype=dt)
        else:
            y = ops.conv([(x, self.kernel * self.kernel)], wt, self.weight.shape[0], bias=bias, pad=self.kernel // 2,
                         out_dtype=dt)
        return y if y.dtype == dt else y.to(dt)
