"""Functional wrappers: torch tensors in, torch tensors out, CUDA work done by libmudiff_b200.

Layout contract: 4-D activations are logical NCHW tensors whose memory is channels-last
(NHWC), i.e. `x.permute(0, 2, 3, 1)` is contiguous.  `as_nhwc()` converts anything else
at the API boundary.  PyTorch is used for allocation, streams and views only.
"""
import ctypes as C
import math

import torch

from . import _lib as L

SQRT2_INV = 1.0 / math.sqrt(2.0)

import os as _os
_ENV_FLAGS = 2 if _os.environ.get('MUDIFF_HALO', '1') == '0' else 0     # debug knob: forbid halo staging
_ENV_FLAGS |= int(_os.environ.get('MUDIFF_XF_DBG', '0')) << 20        # timing ablations of the operand transform
_ENV_FLAGS |= 1024 if _os.environ.get('MUDIFF_BCAP12', '0') == '1' else 0   # ablation: the older, larger B ring (12 sub-tiles)
_ENV_FLAGS |= 0x400000 if _os.environ.get('MUDIFF_NT128', '1') == '0' else 0   # ablation: N = 384 as two 192-column tiles (one accumulator stage) instead of three of 128
_ENV_FLAGS |= int(_os.environ.get('MUDIFF_CONV_DBG_FLAGS', '0'), 0)    # timing ablations only (64: no operand traffic, 128: no epilogue - WRONG RESULTS)

# Fused epilogue statistics (conv_tc butterfly reduction) are implemented and tested, but since the MMA issue
# loop got fast they cost more than the stand-alone HBM-bound statistics pass for every N <= 256 (measured,
# tools/conv_bench.py): off by default, `fused_stats=True` / MUDIFF_FUSED_STATS_MIN_N turn them on.
FUSED_STATS_MIN_N = int(_os.environ.get('MUDIFF_FUSED_STATS_MIN_N', '100000'))

# Optional per-launch profiler (bench.py): callable(kind, flops, bytes) -> context manager or None.
_PROFILER = None


def set_profiler(fn):
    global _PROFILER
    _PROFILER = fn


# ---------------------------------------------------------------------------------
# layout helpers
# ---------------------------------------------------------------------------------
def is_nhwc_view(x: torch.Tensor) -> bool:
    """True if x (logical [B,C,H,W]) is channels-last memory, possibly a channel-slice of a wider
    channels-last tensor (pixel stride ld >= C)."""
    b, c, h, w = x.shape
    if c > 1 and x.stride(1) != 1:
        return False
    if w > 1:
        ld = x.stride(3)
    elif h > 1:
        ld = x.stride(2)
    elif b > 1:
        ld = x.stride(0)
    else:
        return True
    if ld < c:
        return False
    if h > 1 and x.stride(2) != w * ld:
        return False
    if b > 1 and x.stride(0) != h * w * ld:
        return False
    return True


def as_nhwc(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """Return x (logical [B,C,H,W]) with channels-last memory and optional dtype."""
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    if is_nhwc_view(x):
        return x
    return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def empty_nhwc(b, c, h, w, dtype, device) -> torch.Tensor:
    return torch.empty((b, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def _pix_ld(x):
    """pixel stride (elements) of a channels-last tensor or channel-slice view of one."""
    b, c, h, w = x.shape
    if w > 1:
        return x.stride(3)
    if h > 1:
        return x.stride(2)
    if b > 1:
        return x.stride(0)
    return c


def channel_slice(x: torch.Tensor, c0: int, c1: int) -> torch.Tensor:
    return x[:, c0:c1]


# ---------------------------------------------------------------------------------
# FIR / bias-act (the reference's native ops)
# ---------------------------------------------------------------------------------
_kernel_cache = {}


def fir_kernel_device(k_np, device) -> torch.Tensor:
    """fp32 device copy of a (small) FIR kernel, cached so that no H2D copy happens per call
    (the reference rebuilds torch.tensor(k, device=...) on every call,
    up_or_down_sampling.py:145,181,228,261, which is not graph-capture safe)."""
    key = (k_np.tobytes(), k_np.shape, str(device))
    t = _kernel_cache.get(key)
    if t is None:
        t = torch.tensor(k_np, dtype=torch.float32, device=device)
        _kernel_cache[key] = t
    return t


def upfirdn2d_raw(x, kernel_f32, major, in_h, in_w, minor, up, down, pad, out):
    """x/out are flat device buffers in [major,H,W,minor] order."""
    kh, kw = kernel_f32.shape
    rc = L.lib().mudiff_upfirdn2d(x.data_ptr(), out.data_ptr(), kernel_f32.data_ptr(), L.dtype_code(x.dtype),
                                  major, in_h, in_w, minor, kh, kw, up[0], up[1], down[0], down[1],
                                  pad[0], pad[1], pad[2], pad[3], L.stream_ptr(x.device))
    L.check(rc, 'upfirdn2d')


def upfirdn2d_nhwc(x, kernel_f32, up=1, down=1, pad=(0, 0)):
    """FIR on a channels-last activation (all channels share the kernel)."""
    L.require_cuda(x, kernel_f32)
    b, c, h, w = x.shape
    x = as_nhwc(x)
    kh, kw = kernel_f32.shape
    oh = (h * up + pad[0] + pad[1] - kh) // down + 1
    ow = (w * up + pad[0] + pad[1] - kw) // down + 1
    out = empty_nhwc(b, c, oh, ow, x.dtype, x.device)
    if _pix_ld(x) != c:
        x = as_nhwc(x.contiguous(memory_format=torch.channels_last))
    upfirdn2d_raw(x, kernel_f32, b, h, w, c, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]), out)
    return out


FUSED_FIR_GN = _os.environ.get('MUDIFF_FUSED_FIR_GN', '1') != '0'


def fir_resample_gn(x, table, kernel_f32, up: bool, act=L.ACT_SILU):
    """(FIR(act(x * scale + shift)), FIR(x)) of a channels-last activation from ONE read of x: the two resampled
    branches of a resample ResBlock (backbones/layerspp.py:293-305).  `table` = gn_scale_shift() output [B, C, 2];
    `kernel_f32` the 4x4 device FIR kernel with the gain folded in; up=True: upsample_2d (x2), else downsample_2d (/2).
    Returns None when the kernel does not take the shape (the caller then runs GroupNorm-apply + two FIR calls)."""
    if not FUSED_FIR_GN or tuple(kernel_f32.shape) != (4, 4) or x.dtype not in (torch.bfloat16, torch.float32):
        return None
    L.require_cuda(x, table, kernel_f32)
    x = as_nhwc(x)
    b, c, h, w = x.shape
    if _pix_ld(x) != c or c % 4:
        return None
    u, d, p0, p1 = (2, 1, 2, 1) if up else (1, 2, 1, 1)
    oh, ow = (h * u + p0 + p1 - 4) // d + 1, (w * u + p0 + p1 - 4) // d + 1
    out_h = empty_nhwc(b, c, oh, ow, x.dtype, x.device)
    out_x = empty_nhwc(b, c, oh, ow, x.dtype, x.device)
    rc = L.lib().mudiff_upfirdn2d_gn(x.data_ptr(), out_h.data_ptr(), out_x.data_ptr(), kernel_f32.data_ptr(), table.data_ptr(),
                                     table.shape[1], act, L.dtype_code(x.dtype), b, h, w, c, u, d, p0, p1, L.stream_ptr(x.device))
    if rc == L.EUNSUPPORTED:
        return None
    L.check(rc, 'upfirdn2d_gn')
    return out_h, out_x


# ---------------------------------------------------------------------------------
# GroupNorm.  Statistics travel with the tensors as per-channel (sum, sumsq) doubles [B, C, 2]
# (attribute `_mudiff_chstats`), produced by the conv epilogue or by the stand-alone stats kernel.
# ---------------------------------------------------------------------------------
_CHSTATS = '_mudiff_chstats'


def gn_stats(x, out=None):
    """Per-channel (sum, sumsq) of a channels-last tensor -> double [B, C, 2] (or into out=(buf, coff))."""
    x = as_nhwc(x)
    b, c, h, w = x.shape
    if out is None:
        cs = torch.empty((b, c, 2), dtype=torch.float64, device=x.device)
        buf, off = cs, 0
    else:
        buf, off = out
        cs = buf[:, off:off + c]
    rc = L.lib().mudiff_gn_stats(x.data_ptr(), c, _pix_ld(x), L.dtype_code(x.dtype), b, h * w,
                                 buf.data_ptr(), buf.shape[1], off, L.stream_ptr(x.device))
    L.check(rc, 'gn_stats')
    return cs


def set_chstats(x, cs):
    setattr(x, _CHSTATS, cs)
    return x


def get_chstats(x):
    cs = getattr(x, _CHSTATS, None)
    if cs is None:
        cs = gn_stats(x)
        setattr(x, _CHSTATS, cs)
    return cs


def gn_apply(srcs, chstats, groups, gamma=None, beta=None, gb_bstride=0, eps=1e-6, act=L.ACT_NONE,
             out_dtype=None, out=None):
    x0 = srcs[0]
    x1 = srcs[1] if len(srcs) > 1 else None
    s0 = chstats[0]
    s1 = chstats[1] if len(srcs) > 1 else None
    b, c0, h, w = x0.shape
    c = c0 + (x1.shape[1] if x1 is not None else 0)
    out_dtype = out_dtype or x0.dtype
    if out is None:
        out = empty_nhwc(b, c, h, w, out_dtype, x0.device)
    rc = L.lib().mudiff_gn_apply(x0.data_ptr(), c0, _pix_ld(x0), s0.data_ptr(), s0.stride(0) // 2,
                                 x1.data_ptr() if x1 is not None else None,
                                 x1.shape[1] if x1 is not None else 0, _pix_ld(x1) if x1 is not None else 0,
                                 s1.data_ptr() if s1 is not None else None,
                                 s1.stride(0) // 2 if s1 is not None else 0,
                                 L.dtype_code(x0.dtype),
                                 gamma.data_ptr() if gamma is not None else None,
                                 beta.data_ptr() if beta is not None else None, gb_bstride,
                                 out.data_ptr(), _pix_ld(out), L.dtype_code(out.dtype), b, h * w, groups,
                                 float(eps), act, L.stream_ptr(x0.device))
    L.check(rc, 'gn_apply')
    return out


# GroupNorm/AdaGN + SiLU applied by the conv kernel to its staged operand tiles (conv_tc a_xform) instead of a separate
# read + write pass.  Correct and tested, but NOT a win on B200 so far: the in-shared-memory transform adds a read +
# write of every staged tile to shared-memory bandwidth that the N <= 128 tensor-core launches do not have to spare
# (they are bound by the MMA operand reads): conv_tc +46 ms vs 48 ms of gn_apply saved per step when applied
# everywhere, +5 ms vs 3 ms when restricted to single 64-channel segments (profiles/r01_fused_gn_ablation.md).
#   MUDIFF_FUSED_GN = auto (default): where the launch is small enough to be latency-bound rather than shared-memory-bound
#                     (at most XFORM_MAX_PIXELS output pixels: every conv of a B <= 8 batch at 256^2, the 64^2 level of any
#                     batch up to 128) - there the two HBM passes cost more than the transform (B = 1: 22.7 -> 19.3 ms per
#                     sample, B = 4: +6.6 %, profiles/r02_small_batch.md);
#                     0 never, 1 single 64-channel 3x3 segment only, 2 wherever the kernel supports it.
# Fused and unfused paths are BIT-IDENTICAL (same fp32 scale / shift expressions, silu(t) = h + h tanh(h) on h = t / 2 with
# the 1/2 folded into scale and shift exactly; tests/test_gpu_model.py::test_fused_groupnorm_is_bit_identical), so a
# batch-size dependent choice does not break batch invariance.
_fg = _os.environ.get('MUDIFF_FUSED_GN', 'auto')
FUSED_GN = -1 if _fg == 'auto' else int(_fg)
XFORM_MAX_PIXELS = int(_os.environ.get('MUDIFF_XFORM_MAX_PIXELS', str(8 * 65536)))


def xform_profitable(seg_channels, extra_segments=0, pixels=None) -> bool:
    if FUSED_GN >= 2:
        return True
    if FUSED_GN < 0:
        return pixels is not None and pixels <= XFORM_MAX_PIXELS
    return FUSED_GN == 1 and len(seg_channels) == 1 and seg_channels[0] == 64 and extra_segments == 0


FUSED_STATS_TABLE = _os.environ.get('MUDIFF_FUSED_STATS_TABLE', '1') != '0'


def gn_scale_shift(srcs, chstats, groups, gamma=None, beta=None, gb_bstride=0, eps=1e-6):
    """Folded GroupNorm / AdaGN parameters of the channel-concat of `srcs`: float [B, C, 2] = (scale, shift) with
    GN(x)*gamma + beta == x*scale + shift.  Consumed by conv(..., segs=[(x, taps, (table, c_off))]) which applies
    them (and the SiLU) to the staged activation tile instead of a separate full read + write pass."""
    x0 = srcs[0]
    b, c0, h, w = x0.shape
    c1 = srcs[1].shape[1] if len(srcs) > 1 else 0
    table = torch.empty((b, c0 + c1, 2), dtype=torch.float32, device=x0.device)
    if chstats is None:
        # lazily computed statistics: x0's are not known yet -> statistics + table in ONE launch (mudiff_gn_stats_table)
        s1 = getattr(srcs[1], _CHSTATS, None) if len(srcs) > 1 else None
        if FUSED_STATS_TABLE and getattr(x0, _CHSTATS, None) is None and (len(srcs) == 1 or s1 is not None) \
                and x0.dtype in (torch.bfloat16, torch.float32) and is_nhwc_view(x0):
            cs = torch.empty((b, c0, 2), dtype=torch.float64, device=x0.device)
            rc = L.lib().mudiff_gn_stats_table(x0.data_ptr(), c0, _pix_ld(x0), L.dtype_code(x0.dtype), cs.data_ptr(), c0,
                                               c1, s1.data_ptr() if s1 is not None else None,
                                               s1.stride(0) // 2 if s1 is not None else 0,
                                               gamma.data_ptr() if gamma is not None else None,
                                               beta.data_ptr() if beta is not None else None, gb_bstride,
                                               b, h * w, groups, float(eps), table.data_ptr(), L.stream_ptr(x0.device))
            if rc != L.EUNSUPPORTED:
                L.check(rc, 'gn_stats_table')
                set_chstats(x0, cs)
                return table
        chstats = [get_chstats(t) for t in srcs]
    s0 = chstats[0]
    s1 = chstats[1] if len(srcs) > 1 else None
    rc = L.lib().mudiff_gn_scale_shift(s0.data_ptr(), s0.stride(0) // 2, c0,
                                       s1.data_ptr() if s1 is not None else None,
                                       s1.stride(0) // 2 if s1 is not None else 0, c1,
                                       gamma.data_ptr() if gamma is not None else None,
                                       beta.data_ptr() if beta is not None else None, gb_bstride,
                                       b, h * w, groups, float(eps), table.data_ptr(), L.stream_ptr(x0.device))
    L.check(rc, 'gn_scale_shift')
    return table


# Single-pass GroupNorm (statistics + apply with ONE read of the tensor, persistent cooperative kernel, mudiff_gn_fused).
# Correct and tested, but NOT a win on B200 so far: every image costs each CTA a chain of dependent global round trips
# (partial write -> ticket -> cross-CTA reduce -> flag -> statistics -> gamma/beta) that three shared-memory stages do
# not cover: 0.87 ms per launch against 0.30 ms for gn_stats + gn_apply at B = 64, C = 64, 256^2
# (profiles/r01_single_pass_gn.md).  MUDIFF_GN_SINGLE_PASS=1 turns it on.
GN_SINGLE_PASS = _os.environ.get('MUDIFF_GN_SINGLE_PASS', '0') != '0'


def gn_single_pass(x, groups, gamma=None, beta=None, gb_bstride=0, eps=1e-6, act=L.ACT_NONE):
    """act(GN(x)*gamma + beta) AND x's per-channel statistics (attached like get_chstats) from ONE read of x.
    Returns None when the kernel does not take the shape (image chunk too large for the shared-memory stages,
    non-dense or non-bf16 tensor): the caller then runs gn_stats + gn_apply."""
    if not GN_SINGLE_PASS or x.dtype != torch.bfloat16:
        return None
    b, c, h, w = x.shape
    if _pix_ld(x) != c or c % 8:
        return None
    cs = torch.empty((b, c, 2), dtype=torch.float64, device=x.device)
    out = empty_nhwc(b, c, h, w, x.dtype, x.device)
    rc = L.lib().mudiff_gn_fused(x.data_ptr(), out.data_ptr(), c, b, h * w, groups,
                                 gamma.data_ptr() if gamma is not None else None,
                                 beta.data_ptr() if beta is not None else None, gb_bstride, float(eps), act,
                                 cs.data_ptr(), c, 0, L.stream_ptr(x.device))
    if rc == L.EUNSUPPORTED:
        return None
    L.check(rc, 'gn_fused')
    set_chstats(x, cs)
    return out


# Statistics + apply in one launch, the second read of the tensor served by L2 (mudiff_gn_stats_apply); bit-identical to
# gn_stats + gn_apply, so the choice may depend on the batch.  Measured on B200 (profiles/r02_gn_l2.md): at B = 64 the blocks of
# an image wait for each other through a chain of dependent global round trips (partials -> ticket -> totals -> flag) that
# three resident blocks per SM do not cover - 58 ms (256 KB chunks, no L2 reuse) to 106 ms (64 KB chunks) against 43 + 17 ms
# for the two launches - so `auto` (default) takes it only where the tensor fits L2 anyway (B <= 4 at 256^2: one launch less
# per GroupNorm, +0.7 %); 1 = always, 0 = never.
_gl2 = _os.environ.get('MUDIFF_GN_L2', 'auto')
GN_L2 = -1 if _gl2 == 'auto' else int(_gl2)
GN_L2_MAX_BYTES = 48 << 20


def gn_stats_apply(srcs, groups, gamma=None, beta=None, gb_bstride=0, eps=1e-6, act=L.ACT_NONE):
    """act(GN([x0 | x1]) * gamma + beta) where x0's statistics are not known yet (computed here and attached to x0) and the
    optional x1 already carries its own.  Returns None when the kernel does not take the shape."""
    x0 = srcs[0]
    x1 = srcs[1] if len(srcs) > 1 else None
    if not GN_L2 or x0.dtype != torch.bfloat16 or (x1 is not None and x1.dtype != torch.bfloat16):
        return None
    b, c0, h, w = x0.shape
    if GN_L2 < 0 and x0.numel() * 2 > GN_L2_MAX_BYTES:
        return None
    c1 = x1.shape[1] if x1 is not None else 0
    s1 = getattr(x1, _CHSTATS) if x1 is not None else None
    cs = torch.empty((b, c0, 2), dtype=torch.float64, device=x0.device)
    out = empty_nhwc(b, c0 + c1, h, w, x0.dtype, x0.device)
    rc = L.lib().mudiff_gn_stats_apply(x0.data_ptr(), c0, _pix_ld(x0), cs.data_ptr(), c0,
                                       x1.data_ptr() if x1 is not None else None, c1, _pix_ld(x1) if x1 is not None else 0,
                                       s1.data_ptr() if s1 is not None else None, s1.stride(0) // 2 if s1 is not None else 0,
                                       L.dtype_code(x0.dtype),
                                       gamma.data_ptr() if gamma is not None else None,
                                       beta.data_ptr() if beta is not None else None, gb_bstride,
                                       out.data_ptr(), c0 + c1, b, h * w, groups, float(eps), act, L.stream_ptr(x0.device))
    if rc == L.EUNSUPPORTED:
        return None
    L.check(rc, 'gn_stats_apply')
    set_chstats(x0, cs)
    return out


def gn_apply_auto(srcs, groups, gamma=None, beta=None, gb_bstride=0, eps=1e-6, act=L.ACT_NONE):
    """GroupNorm (+AdaGN, +act) of one tensor or the channel-concat of two: one launch (statistics + apply, gn_stats_apply)
    when the statistics of the first source are not known yet, else statistics (cached per tensor) + apply."""
    if getattr(srcs[0], _CHSTATS, None) is None and (len(srcs) == 1 or getattr(srcs[1], _CHSTATS, None) is not None):
        if len(srcs) == 1:
            y = gn_single_pass(srcs[0], groups, gamma, beta, gb_bstride, eps, act)
            if y is not None:
                return y
        y = gn_stats_apply(srcs, groups, gamma, beta, gb_bstride, eps, act)
        if y is not None:
            return y
    return gn_apply(srcs, [get_chstats(t) for t in srcs], groups, gamma=gamma, beta=beta, gb_bstride=gb_bstride, eps=eps, act=act)


def group_norm(srcs, groups, gamma=None, beta=None, gb_bstride=0, eps=1e-6, act=L.ACT_NONE, out_dtype=None):
    srcs = [s if is_nhwc_view(s) else as_nhwc(s) for s in srcs]
    if out_dtype is None or out_dtype == srcs[0].dtype:
        return gn_apply_auto(srcs, groups, gamma, beta, gb_bstride, eps, act)
    return gn_apply(srcs, [get_chstats(s) for s in srcs], groups, gamma, beta, gb_bstride, eps, act, out_dtype)


# ---------------------------------------------------------------------------------
# convolution / contraction
# ---------------------------------------------------------------------------------
def pack_conv_weight(weight: torch.Tensor, seg_channels, dtype) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, sum_seg kh*kw*C_seg] with k = seg_off + tap*C_seg + c."""
    parts, off = [], 0
    cout = weight.shape[0]
    for cs in seg_channels:
        parts.append(weight[:, off:off + cs].permute(0, 2, 3, 1).reshape(cout, -1))
        off += cs
    assert off == weight.shape[1], (off, weight.shape)
    return torch.cat(parts, dim=1).to(dtype).contiguous()


# fp32 parity path on the tensor cores: every fp32 operand is split into three bf16 terms (x = hi + mid + lo exactly) and the
# conv runs on the tcgen05 kernel as the six bf16 products above 2^-24, accumulated in fp32 (TMEM):
#   [x_hi] * w_lo  +  [x_mid x_hi] * w_mid  +  [x_lo x_mid x_hi] * w_hi        (three channel-slice segments of ONE tensor
#   [lo | mid | hi]; small products FIRST: the tensor core's fp32 accumulation truncates relative to the accumulator's
#   magnitude, so the 2^-16 / 2^-8 terms are added while it is still small and hi x hi comes last)
# 6x the bf16 FLOPs, still ~6x faster than the CUDA-core fp32 implicit GEMM.  MUDIFF_FP32_TC=0 restores the CUDA-core path.
FP32_TC = _os.environ.get('MUDIFF_FP32_TC', '1') != '0'
_split_w_cache = {}


def split3(x):
    """fp32 channels-last [B, C, H, W] -> bf16 [B, 3C, H, W] = (lo | mid | hi), x == hi + mid + lo."""
    x = as_nhwc(x)
    b, c, h, w = x.shape
    out = empty_nhwc(b, 3 * c, h, w, torch.bfloat16, x.device)
    L.check(L.lib().mudiff_split3_bf16(x.data_ptr(), _pix_ld(x), out.data_ptr(), b * h * w, c, 0, L.stream_ptr(x.device)), 'split3_bf16')
    return out


def split3_rows(wt, rows, k, ld):
    """The K-major second operand when it is an fp32 ACTIVATION (`rows` rows of k values, row stride ld, e.g. the keys of
    the attention scores): bf16 [rows, 6k] = (hi hi hi | mid mid | lo), matching the three segments of split3()."""
    out = torch.empty((rows, 6 * k), dtype=torch.bfloat16, device=wt.device)
    L.check(L.lib().mudiff_split3_bf16(wt.data_ptr(), ld, out.data_ptr(), rows, k, 1, L.stream_ptr(wt.device)), 'split3_bf16')
    return out


def _split3_weights_build(wt, c, taps):
    n = wt.shape[0]
    w = wt.detach().view(n, taps, c)
    hi = w.to(torch.bfloat16)
    r1 = w - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    return torch.cat([lo.reshape(n, -1), mid.unsqueeze(2).expand(n, taps, 2, c).reshape(n, -1),
                      hi.unsqueeze(2).expand(n, taps, 3, c).reshape(n, -1)], dim=1).contiguous()


def split3_weights(wt, c, taps):
    """bf16 [N, taps * 6C] for the three segments above, cached per packed fp32 weight tensor; a changed weight is
    re-split INTO the cached tensor (its address stays valid for captured graphs, see refresh_split_weights)."""
    key = (wt.data_ptr(), tuple(wt.shape), taps, wt.device)
    hit = _split_w_cache.get(key)
    if hit is not None and hit[0] == wt._version:
        return hit[1]
    with torch.no_grad():
        val = _split3_weights_build(wt, c, taps)
        if hit is not None:
            hit[1].copy_(val)
            val = hit[1]
    _split_w_cache[key] = (wt._version, val, wt, c)
    return val


def refresh_split_weights():
    for key, (ver, val, wt, c) in list(_split_w_cache.items()):
        if wt._version != ver:
            split3_weights(wt, c, key[2])


def fp32_tc_eligible(segs, n, stride, wt, w_bstride, w_ld, a_batched, batch, dec2) -> bool:
    """One fp32 segment, stride 1, Cin % 64 == 0, N % 32 == 0.  Packed weights (w_bstride == 0) are split once and cached;
    a per-sample / strided second operand (attention: w_bstride = N * w_ld) is split per call and must be 1-tap."""
    if not FP32_TC or len(segs) != 1 or len(segs[0]) > 2 or stride != 1 or dec2 or n % 32:
        return False
    x, taps = segs[0][0], segs[0][1]
    if x.dtype != torch.float32 or wt.dtype != torch.float32 or x.shape[1] % 64:
        return False
    k = taps * x.shape[1]
    if w_bstride == 0 and w_ld == 0:
        return a_batched and batch is None and wt.is_contiguous() and wt.ndim == 2 and wt.shape[1] == k
    ld = w_ld or k
    return taps == 1 and ld % 4 == 0 and (w_bstride == 0 or w_bstride == n * ld)


def tc_eligible(segs, n, stride, dtype) -> bool:
    if dtype != torch.bfloat16 or stride != 1 or n % 32:
        return False
    return all(sg[0].shape[1] % 64 == 0 for sg in segs)


def conv(segs, wt, n, *, bias=None, rowbias=None, residual=None, alpha=1.0, beta=0.0, act=L.ACT_NONE,
         out=None, out_coff=0, out_dtype=None, stride=1, pad=1, w_bstride=0, w_ld=0, a_batched=True,
         batch=None, flags=0, force=None, want_stats=False, stats_out=None, dec2=False, fused_stats=None,
         xform_act=L.ACT_SILU):
    """Implicit-GEMM convolution.  segs = [(tensor NCHW-logical/channels-last, taps[, (table, c_off)])], wt packed
    K-major; the optional third entry is a gn_scale_shift() table: the segment is read as
    xform_act(x * scale + shift) (GroupNorm/AdaGN + SiLU fused into the operand path, tcgen05 kernel only).
    Chooses the tcgen05 kernel when eligible (bf16, Cin % 64 == 0, N % 32 == 0, stride 1), otherwise the
    CUDA-core kernel.  `force` in {None,'tc','simt'}."""
    x0 = segs[0][0]
    if force is None and fp32_tc_eligible(segs, n, stride, wt, w_bstride, w_ld, a_batched, batch, dec2):
        c, taps = x0.shape[1], segs[0][1]
        x3 = split3(x0)
        if w_bstride == 0 and w_ld == 0:
            w3, wb3, wl3 = split3_weights(wt, c, taps), 0, 0
        else:                                   # the second operand is an activation (attention): split it per call
            nb = (batch if batch is not None else x0.shape[0]) if w_bstride else 1
            w3 = split3_rows(wt, nb * n, c, w_ld or c)
            wl3, wb3 = 6 * c, (n * 6 * c if w_bstride else 0)
        return conv([(x3[:, 2 * c:], taps), (x3[:, c:], taps), (x3, taps)], w3, n, bias=bias,
                    rowbias=rowbias, residual=residual, alpha=alpha, beta=beta, act=act, out=out, out_coff=out_coff,
                    out_dtype=out_dtype or torch.float32, pad=pad, w_bstride=wb3, w_ld=wl3, a_batched=a_batched, batch=batch,
                    flags=flags, force='tc', want_stats=want_stats, stats_out=stats_out, fused_stats=fused_stats)
    dev = x0.device
    b = batch if batch is not None else x0.shape[0]
    h, w = x0.shape[2], x0.shape[3]
    if dec2:                      # stride-2 VALID 3x3 conv == odd outputs of the pad-1 'same' conv (tensor-core path)
        ho, wo = (h - 1) // 2, (w - 1) // 2
        flags |= 0x8000
    elif stride == 1:
        ho, wo = h, w
    else:
        k = 3 if any(t == 9 for _, t in segs) else 1
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out_dtype = out_dtype or x0.dtype
    fresh = out is None
    if out is None:
        out = empty_nhwc(b, n, ho, wo, out_dtype, dev)
    d = L.ConvDesc()
    keep = []
    any_xform = False
    for i, sg in enumerate(segs):
        t, taps = sg[0], sg[1]
        t = as_nhwc(t)
        keep.append(t)
        d.a[i] = t.data_ptr()
        d.a_c[i] = t.shape[1]
        d.a_ld[i] = _pix_ld(t)
        d.a_taps[i] = taps
        if len(sg) > 2 and sg[2] is not None:
            table, coff = sg[2]
            keep.append(table)
            d.a_xform[i] = table.data_ptr() + coff * 8
            d.a_xform_ld[i] = table.shape[1]
            any_xform = True
    d.a_xform_act = xform_act if any_xform else L.ACT_NONE
    d.nseg = len(segs)
    d.a_batched = 1 if a_batched else 0
    d.batch, d.h, d.w = b, h, w
    d.stride, d.pad = stride, pad
    d.wt = wt.data_ptr()
    d.w_bstride = w_bstride
    d.w_ld = w_ld
    d.n = n
    d.bias = bias.data_ptr() if bias is not None else None
    d.rowbias = rowbias.data_ptr() if rowbias is not None else None
    d.rowbias_ld = rowbias.stride(0) if rowbias is not None else 0
    if residual is not None:
        residual = as_nhwc(residual)
        if residual.dtype != out.dtype:
            residual = residual.to(out.dtype)
        d.residual = residual.data_ptr()
        d.res_ld = _pix_ld(residual)
    d.alpha, d.beta, d.act = float(alpha), float(beta), act
    d.out = out.data_ptr()
    d.out_ld = _pix_ld(out)
    d.out_coff = out_coff
    d.out_dtype = L.dtype_code(out.dtype)
    d.stats = None
    d.stats_groups = 0
    d.flags = flags | _ENV_FLAGS
    use_tc = tc_eligible(segs, n, stride, x0.dtype) and wt.dtype == torch.bfloat16
    if force == 'tc':
        use_tc = True
    elif force == 'simt':
        use_tc = False
    st = L.stream_ptr(dev)
    partial = None
    tpi = 0
    # fused epilogue statistics pay off only when the MMA phase of a tile is long enough to hide the
    # butterfly reduction (N >= 128); N = 64 outputs get a stand-alone statistics pass (HBM-bound, cheaper)
    use_fused = (n >= FUSED_STATS_MIN_N) if fused_stats is None else bool(fused_stats)
    if want_stats and use_tc and use_fused and n <= 256:
        q = (C.c_int32 * 10)()
        L.check(L.lib().mudiff_conv_tc_query(C.byref(d), q), 'conv_tc_query')
        tpi = q[2]
        partial = torch.empty((b * tpi * 4, n, 2), dtype=torch.float32, device=dev)   # row = (image, tile, lane quadrant)
        d.stats = partial.data_ptr()
    prof = None
    if _PROFILER is not None:
        ktot = sum(sg[0].shape[1] * sg[1] for sg in segs)
        prof = _PROFILER('conv_tc' if use_tc else 'conv_simt', 2.0 * b * ho * wo * n * ktot,
                         dict(n=n, ktot=ktot, pixels=b * ho * wo))
        prof.__enter__()
    if any_xform and not use_tc:
        raise RuntimeError("mu-diff_b200: the fused GroupNorm operand transform needs the tcgen05 conv kernel")
    if use_tc:
        L.check(L.lib().mudiff_conv_tc(C.byref(d), st), 'conv_tc')
    else:
        if wt.dtype != x0.dtype:
            raise RuntimeError("mu-diff_b200: conv_simt needs weights in the activation dtype")
        L.check(L.lib().mudiff_conv_simt(C.byref(d), L.dtype_code(x0.dtype), st), 'conv_simt')
    if prof is not None:
        prof.__exit__(None, None, None)
    if want_stats:
        region = out if (fresh or (out_coff == 0 and out.shape[1] == n)) else out[:, out_coff:out_coff + n]
        if partial is not None:
            if stats_out is None:
                cs = torch.empty((b, n, 2), dtype=torch.float64, device=dev)
                buf, off = cs, 0
            else:
                buf, off = stats_out
                cs = buf[:, off:off + n]
            L.check(L.lib().mudiff_stats_finalize(partial.data_ptr(), tpi * 4, n, buf.data_ptr(), buf.shape[1], off, b, st),
                    'stats_finalize')
        elif stats_out is not None or not ((GN_SINGLE_PASS or GN_L2 or FUSED_STATS_TABLE) and region is out):
            cs = gn_stats(region, out=stats_out)
        else:
            cs = None                    # lazy: the consuming GroupNorm computes them (single-pass kernel or get_chstats)
        if region is out and cs is not None:
            set_chstats(out, cs)
    return out


FUSED_STEM = _os.environ.get('MUDIFF_FUSED_STEM', '1') != '0'


_MOM_SCOPE = None


class stem_moments_scope:
    """Inside the scope (one pass of the sampling loop, sampling.sample_from_model / GraphSampler) the second moments of a
    1-channel input image are computed ONCE per distinct tensor: the three conditioning contrasts feed a stem in each of the
    8 generator forwards of a sample and x_t feeds one in G1 and in G2, so 36 mudiff_stem_moments launches per sample become 11
    (3 contrasts + x_t and x_0' of each of the 4 steps).  The moments depend on the image only (not on the stem's weights), so
    the values are identical.  Keyed by (address, shape, strides, version) with the tensor kept alive until the scope ends;
    nothing inside the loop writes these tensors in place.  Outside a scope nothing is cached."""

    def __enter__(self):
        global _MOM_SCOPE
        self.prev = _MOM_SCOPE
        if _MOM_SCOPE is None:
            _MOM_SCOPE = {}
        return self

    def __exit__(self, *exc):
        global _MOM_SCOPE
        _MOM_SCOPE = self.prev
        return False


def loop_scope():
    """The dict of the active stem_moments_scope (one pass of the sampling loop), or None.  Also used by the generators to keep
    what does not change between the diffusion steps of a sample (_generator._stem)."""
    return _MOM_SCOPE


def stem_conv_gn_act(x, wt9, bias, groups, *, gamma=None, beta=None, gb_bstride=0, eps=1e-6, act=L.ACT_SILU,
                     out_dtype=torch.bfloat16):
    """act(GroupNorm(conv3x3(x))) for a 1-channel input (the stems of ConvFeatBlock / ConvBlock / ConvBlock_GAP,
    backbones/layerspp.py:394-501) without storing the raw conv output: the GroupNorm statistics come from second
    moments of the input image (mudiff_stem_moments), the conv kernel applies the folded scale / shift / activation.
    wt9 = fp32 packed weights [n, 9]."""
    L.require_cuda(x, wt9)
    x = as_nhwc(x, torch.float32)
    b, c, h, w = x.shape
    if c != 1:
        raise RuntimeError("mu-diff_b200: stem_conv_gn_act expects a 1-channel input")
    n = wt9.shape[0]
    dev = x.device
    st = L.stream_ptr(dev)
    key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x._version)
    ent = _MOM_SCOPE.get(key) if _MOM_SCOPE is not None else None
    if ent is None:
        mom = torch.empty((b, 54), dtype=torch.float64, device=dev)
        L.check(L.lib().mudiff_stem_moments(x.data_ptr(), _pix_ld(x), b, h, w, mom.data_ptr(), st), 'stem_moments')
        if _MOM_SCOPE is not None:
            _MOM_SCOPE[key] = (mom, x)             # x is kept alive: its address cannot be reused inside the scope
    else:
        mom = ent[0]
    ss = torch.empty((b, n, 2), dtype=torch.float32, device=dev)
    out = empty_nhwc(b, n, h, w, out_dtype, dev)
    rc = L.lib().mudiff_stem_conv_gn_act(x.data_ptr(), _pix_ld(x), wt9.data_ptr(),
                                         bias.data_ptr() if bias is not None else None, mom.data_ptr(),
                                         gamma.data_ptr() if gamma is not None else None,
                                         beta.data_ptr() if beta is not None else None, gb_bstride, groups, float(eps),
                                         act, ss.data_ptr(), out.data_ptr(), _pix_ld(out), 0, L.dtype_code(out.dtype),
                                         b, h, w, n, st)
    L.check(rc, 'stem_conv_gn_act')
    return out


FUSED_ATTENTION = _os.environ.get('MUDIFF_FUSED_ATTN', '1') != '0'


def attention_supported(c, lt, dtype) -> bool:
    """Shapes the fused tcgen05 attention kernel takes (others use the unfused GEMM / softmax / GEMM kernels)."""
    return FUSED_ATTENTION and dtype == torch.bfloat16 and c == 256 and lt % 128 == 0


def attention(qk, vt, b, lt, c, scale):
    """out[b] = softmax(Q K^T * scale) V with qk = [B, L, 2C] (q | k), vt = V^T [B, C, L] -> [B, L, C] (bf16).
    One kernel: the [L, L] scores never leave the SM (backbones/layerspp.py:118-122 materialises them)."""
    L.require_cuda(qk, vt)
    out = torch.empty((b, lt, c), dtype=torch.bfloat16, device=qk.device)
    rc = L.lib().mudiff_attention_tc(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), b, lt, c, float(scale),
                                     L.stream_ptr(qk.device))
    L.check(rc, 'attention_tc')
    return out


# ---------------------------------------------------------------------------------
# small dense / embeddings
# ---------------------------------------------------------------------------------
def linear(x, weight, bias=None, act_in=L.ACT_NONE, act_out=L.ACT_NONE, out=None):
    """fp32 [B,K] x [J,K]^T (+bias) -> [B,J]."""
    x = x.float().contiguous() if (x.dtype != torch.float32 or x.stride(-1) != 1) else x
    b, k = x.shape
    j = weight.shape[0]
    if out is None:
        out = torch.empty((b, j), dtype=torch.float32, device=x.device)
    rc = L.lib().mudiff_linear(x.data_ptr(), x.stride(0), weight.data_ptr(),
                               bias.data_ptr() if bias is not None else None, out.data_ptr(), out.stride(0),
                               b, k, j, act_in, act_out, L.stream_ptr(x.device))
    L.check(rc, 'linear')
    return out


def timestep_embedding(t, dim, max_positions=10000.0):
    t = t.to(torch.int64).contiguous()
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    L.check(L.lib().mudiff_timestep_embedding(t.data_ptr(), out.data_ptr(), t.shape[0], dim, float(max_positions),
                                              L.stream_ptr(t.device)), 'timestep_embedding')
    return out


def pixelnorm(z):
    z = z.float().contiguous()
    out = torch.empty_like(z)
    L.check(L.lib().mudiff_pixelnorm(z.data_ptr(), out.data_ptr(), z.shape[0], z.shape[1], L.stream_ptr(z.device)), 'pixelnorm')
    return out


# ---------------------------------------------------------------------------------
# elementwise glue
# ---------------------------------------------------------------------------------
def gate_mul(a, b):
    a, b = as_nhwc(a), as_nhwc(b)
    bb, c, h, w = a.shape
    out = empty_nhwc(bb, c, h, w, a.dtype, a.device)
    L.check(L.lib().mudiff_gate_mul(a.data_ptr(), _pix_ld(a), b.data_ptr(), _pix_ld(b), out.data_ptr(), _pix_ld(out),
                                    L.dtype_code(a.dtype), bb * h * w, c, L.stream_ptr(a.device)), 'gate_mul')
    return out


def gate_blend(g, a, b, out=None):
    g, a, b = as_nhwc(g), as_nhwc(a), as_nhwc(b)
    bb, c, h, w = a.shape
    if out is None:
        out = empty_nhwc(bb, c, h, w, a.dtype, a.device)
    L.check(L.lib().mudiff_gate_blend(g.data_ptr(), _pix_ld(g), a.data_ptr(), _pix_ld(a), b.data_ptr(), _pix_ld(b),
                                      out.data_ptr(), _pix_ld(out), L.dtype_code(a.dtype), bb * h * w, c,
                                      L.stream_ptr(a.device)), 'gate_blend')
    return out


def add_scale(a, b, scale):
    a, b = as_nhwc(a), as_nhwc(b)
    if _pix_ld(a) != a.shape[1] or _pix_ld(b) != b.shape[1]:
        raise RuntimeError("mu-diff_b200: add_scale needs dense tensors")
    out = empty_nhwc(*a.shape[:1], a.shape[1], a.shape[2], a.shape[3], a.dtype, a.device)
    L.check(L.lib().mudiff_add_scale(a.data_ptr(), b.data_ptr(), out.data_ptr(), L.dtype_code(a.dtype), a.numel(),
                                     float(scale), L.stream_ptr(a.device)), 'add_scale')
    return out


def copy_channels(src, dst, coff=0):
    """dst[:, coff:coff+C] = src (dtype conversion allowed)."""
    src = as_nhwc(src)
    b, c, h, w = src.shape
    dview = dst[:, coff:coff + c]
    L.check(L.lib().mudiff_copy_channels(src.data_ptr(), _pix_ld(src), L.dtype_code(src.dtype), dview.data_ptr(),
                                         _pix_ld(dst), L.dtype_code(dst.dtype), b * h * w, c,
                                         L.stream_ptr(src.device)), 'copy_channels')
    return dst


def concat(tensors, dtype=None):
    tensors = [as_nhwc(t) for t in tensors]
    b, _, h, w = tensors[0].shape
    c = sum(t.shape[1] for t in tensors)
    out = empty_nhwc(b, c, h, w, dtype or tensors[0].dtype, tensors[0].device)
    off = 0
    for t in tensors:
        copy_channels(t, out, off)
        off += t.shape[1]
    return out


def gap(x):
    x = as_nhwc(x)
    b, c, h, w = x.shape
    out = torch.empty((b, c), dtype=torch.float32, device=x.device)
    L.check(L.lib().mudiff_gap(x.data_ptr(), _pix_ld(x), L.dtype_code(x.dtype), out.data_ptr(), b, h * w, c,
                               L.stream_ptr(x.device)), 'gap')
    return out


def softmax_rows_(x2d, scale=1.0):
    rows, cols = x2d.shape
    L.check(L.lib().mudiff_softmax_rows(x2d.data_ptr(), x2d.data_ptr(), L.dtype_code(x2d.dtype), rows, cols,
                                        float(scale), L.stream_ptr(x2d.device)), 'softmax_rows')
    return x2d


def leaky_relu(x, slope=0.2):
    """LeakyReLU(slope) of a dense tensor (any layout; `mudiff_fused_bias_act` without bias, scale 1)."""
    L.require_cuda(x)
    if not (x.is_contiguous() or (x.ndim == 4 and is_nhwc_view(x) and _pix_ld(x) == x.shape[1])):
        x = as_nhwc(x.contiguous())
    out = torch.empty_like(x)
    L.check(L.lib().mudiff_fused_bias_act(x.data_ptr(), None, None, out.data_ptr(), L.dtype_code(x.dtype), x.numel(), 1, 1,
                                          3, 0, float(slope), 1.0, L.stream_ptr(x.device)), 'fused_bias_act')
    return out


def minibatch_stddev(x, out, out_c, group):
    """Writes the minibatch-stddev feature of x (backbones/discriminator.py:243-250) into channel `out_c` of `out`."""
    x = as_nhwc(x)
    b, c, h, w = x.shape
    xin = x if x.dtype == out.dtype else x.to(out.dtype)
    L.check(L.lib().mudiff_minibatch_stddev(xin.data_ptr(), _pix_ld(xin), out.data_ptr(), _pix_ld(out), out_c,
                                            L.dtype_code(out.dtype), b, group, c, h * w, L.stream_ptr(x.device)),
            'minibatch_stddev')
    return out


def posterior_update(x01, x02, xt, noise, t, coef1, coef2, logvar):
    """engine/test.py:150-177 as one kernel (fp32)."""
    L.require_cuda(x01, x02, xt, noise, t)
    b = xt.shape[0]
    per = xt[0].numel()
    xt = xt.float().contiguous()
    noise = noise.float().contiguous()

    def prep(x):
        x = x.float()
        if x[0].numel() != per:
            raise RuntimeError("mu-diff_b200: posterior_update shape mismatch")
        if not x[0].is_contiguous():
            x = x.contiguous()
        return x, (x.stride(0) if b > 1 else per)

    x01, s1 = prep(x01)
    x02, s2 = prep(x02)
    t = t.to(torch.int64).contiguous()
    out = torch.empty_like(xt)
    rc = L.lib().mudiff_posterior_update(x01.data_ptr(), s1, x02.data_ptr(), s2, xt.data_ptr(), noise.data_ptr(),
                                         t.data_ptr(), coef1.data_ptr(), coef2.data_ptr(), logvar.data_ptr(),
                                         coef1.numel(), out.data_ptr(), b, per, L.stream_ptr(xt.device))
    L.check(rc, 'posterior_update')
    return out
