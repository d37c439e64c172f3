"""`fused_leaky_relu` / `FusedLeakyReLU` with the reference signatures
(utils/op/fused_act.py:100-123), executed by `mudiff_fused_bias_act`.
Not on the generator path - kept because the operator API is part of
the drop-in boundary.  CUDA only (no CPU branch; note the reference's CPU branch
ignores `negative_slope`, :117 - the CUDA semantics are what is implemented)."""
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib as L


def _bias_act(x, bias, ref, act, grad, alpha, scale):
    if not x.is_cuda:
        raise RuntimeError("fused_leaky_relu: expected a CUDA tensor (mu-diff_b200 has no CPU path)")
    x = x.contiguous()
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    step_b = 1
    size_b = 1
    if bias is not None and bias.numel() > 0:
        bias = bias.to(x.dtype).contiguous()
        size_b = bias.shape[0]
        for d in x.shape[2:]:
            step_b *= d                      # fused_bias_act_kernel.cu:70-73 (x.stride(1) of contiguous NCHW)
    else:
        bias = None
    if ref is not None and ref.numel() > 0:
        ref = ref.to(x.dtype).contiguous()
    else:
        ref = None
    rc = L.lib().mudiff_fused_bias_act(x.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       ref.data_ptr() if ref is not None else None, out.data_ptr(),
                                       L.dtype_code(x.dtype), x.numel(), size_b, step_b, act, grad,
                                       float(alpha), float(scale), L.stream_ptr(x.device))
    L.check(rc, 'fused_bias_act')
    return out


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = _bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        dim = [0] + list(range(2, grad_input.ndim))
        return grad_input, grad_input.sum(dim).detach()

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        return _bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale), None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = _bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        gi, gb = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return gi, gb, None, None


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
