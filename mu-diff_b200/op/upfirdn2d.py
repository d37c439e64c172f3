"""`upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))` - same signature, argument meaning
and error behaviour as utils/op/upfirdn2d.py:170-199 of the reference, executed by the
sm_100a kernels of libmudiff_b200 (C ABI `mudiff_upfirdn2d`).

Differences by design: there is NO CPU path (the reference routes CPU tensors to
`upfirdn2d_native`, :171-174); a CPU tensor raises RuntimeError here.  Autograd is wired
exactly like the reference (the gradient of upfirdn2d is an upfirdn2d with swapped
up/down, flipped kernel and the g_pad of :135-140).
"""
from collections import abc

import torch
from torch.autograd import Function

from .. import _lib as L
from .. import ops


def _run(inp, kernel, up, down, pad):
    """inp [N,C,H,W] (any memory format), kernel [kh,kw] -> [N,C,oh,ow] in inp's memory format."""
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    n, c, h, w = inp.shape
    kh, kw = kernel.shape
    oh = (h * up_y + py0 + py1 - kh) // down_y + 1
    ow = (w * up_x + px0 + px1 - kw) // down_x + 1
    if oh < 1 or ow < 1:
        raise RuntimeError("upfirdn2d: output would be empty (kernel larger than padded input)")
    k32 = kernel.detach().to(torch.float32).contiguous()
    chlast = c > 1 and inp.stride(1) == 1 and inp.permute(0, 2, 3, 1).is_contiguous()
    if chlast:       # NHWC: major = N, minor = C
        out = ops.empty_nhwc(n, c, oh, ow, inp.dtype, inp.device)
        ops.upfirdn2d_raw(inp, k32, n, h, w, c, up, down, pad, out)
    else:            # NCHW planes: major = N*C, minor = 1   (reference: input.reshape(-1, H, W, 1), :124)
        x = inp.contiguous()
        out = torch.empty((n, c, oh, ow), dtype=inp.dtype, device=inp.device)
        ops.upfirdn2d_raw(x, k32, n * c, h, w, 1, up, down, pad, out)
    return out


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        grad_input = _run(grad_output.reshape(in_size[0], in_size[1], out_size[0], out_size[1]),
                          grad_kernel, down, up, g_pad)
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad, ctx.in_size, ctx.out_size = up, down, pad, in_size, out_size
        return grad_input.reshape(in_size)

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gg = _run(gradgrad_input.reshape(ctx.in_size), kernel, ctx.up, ctx.down, ctx.pad)
        return gg, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = tuple(input.shape)
        out = _run(input, kernel, up, down, pad)
        out_h, out_w = out.shape[2], out.shape[3]
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.g_pad = (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                     kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        gi = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                     ctx.in_size, ctx.out_size)
        return gi, None, None, None, None


def _check(input, kernel):
    if input.ndim != 4 or kernel.ndim != 2:
        raise RuntimeError("upfirdn2d: input must be [N,C,H,W] and kernel [kh,kw]")
    if not input.is_cuda or not kernel.is_cuda:
        # reference: TORCH_CHECK(is_cuda) -> RuntimeError (utils/op/upfirdn2d.cpp:16,23-24)
        raise RuntimeError("upfirdn2d: input and kernel must be CUDA tensors (mu-diff_b200 has no CPU path)")


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    _check(input, kernel)
    if input.numel() == 0:
        kh, kw = kernel.shape
        n, c, h, w = input.shape
        return input.new_empty((n, c, (h * up + pad[0] + pad[1] - kh) // down + 1, (w * up + pad[0] + pad[1] - kw) // down + 1))
    args = (input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
    if input.requires_grad and torch.is_grad_enabled():
        return UpFirDn2d.apply(*args)
    return _run(*args)


def upfirdn2d_ada(input, kernel, up=1, down=1, pad=(0, 0)):
    _check(input, kernel)
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    args = (input, kernel, tuple(up), tuple(down), tuple(pad))
    if input.requires_grad and torch.is_grad_enabled():
        return UpFirDn2d.apply(*args)
    return _run(*args)
