"""Drop-in for the reference's ``models.StyleGAN2.op`` package (op/__init__.py:1-2): ``FusedLeakyReLU``,
``fused_leaky_relu`` and ``upfirdn2d`` with the reference's signatures, plus the two raw extension functions the
reference binds with pybind11 (``fused_bias_act``, ``upfirdn2d_op``), all over libwgs_b200 (csrc/ops.cu).

First-order autograd only (the reference also defines double-backward, op/fused_act.py:41-48 and
op/upfirdn2d.py:61-84; nothing on the WarpedGANSpace path differentiates twice).  CUDA fp32 tensors only - like the
reference's extensions, which CHECK_CUDA (op/fused_bias_act.cpp:13-14), there is no CPU path.
"""
import torch
from torch import nn

from .. import _lib

__all__ = ['FusedLeakyReLU', 'fused_leaky_relu', 'upfirdn2d', 'fused_bias_act', 'upfirdn2d_op']


def _check(t, name):
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor' % name)              # the reference's CHECK_CUDA message
    if t.dtype != torch.float32:
        raise RuntimeError('%s: libwgs_b200 ops are fp32 (got %s)' % (name, t.dtype))


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    """``fused.fused_bias_act`` (op/fused_bias_act.cpp:11-20): empty `bias` / `refer` tensors mean "absent"."""
    _check(input, 'input')
    x = input.contiguous()
    b = bias.contiguous() if bias.numel() else None
    r = refer.contiguous() if refer.numel() else None
    out = torch.empty_like(x)
    step_b = 1
    for s in x.shape[2:]:
        step_b *= s
    size_b = x.shape[1] if x.dim() > 1 else 1
    _lib.call('wgs_fused_bias_act', _lib.ptr(x), _lib.ptr(b), _lib.ptr(r), _lib.ptr(out), x.numel(), step_b, size_b,
              int(act), int(grad), float(alpha), float(scale), _lib.stream())
    return out


def upfirdn2d_op(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """``upfirdn2d_op.upfirdn2d`` (op/upfirdn2d.cpp:12-22): input [major, in_h, in_w, minor] -> [major, out_h, out_w, minor]."""
    _check(input, 'input')
    _check(kernel, 'kernel')
    x, k = input.contiguous(), kernel.contiguous()
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) // down_y + 1
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) // down_x + 1
    out = x.new_empty(major, max(out_h, 0), max(out_w, 0), minor)
    _lib.call('wgs_upfirdn2d', _lib.ptr(x), _lib.ptr(k), _lib.ptr(out), major, in_h, in_w, minor, kh, kw, up_x, up_y,
              down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, _lib.stream())
    return out


class _FusedLeakyReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = fused_bias_act(input, bias, input.new_empty(0), 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input = fused_bias_act(grad_output, grad_output.new_empty(0), out, 3, 1, ctx.negative_slope, ctx.scale)
        dims = [0] + list(range(2, grad_input.ndim))                  # op/fused_act.py:32-37
        return grad_input, grad_input.sum(dims), None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:87-88: ``scale * leaky_relu(input + bias[None, :, None, ...], negative_slope)``."""
    return _FusedLeakyReLUFn.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    """op/fused_act.py:73-84."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


class _UpFirDn2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kh, kw = kernel.shape
        n, c, in_h, in_w = input.shape
        out = upfirdn2d_op(input.reshape(-1, in_h, in_w, 1), kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        out_h, out_w = out.shape[1], out.shape[2]
        ctx.save_for_backward(torch.flip(kernel, [0, 1]))
        # the adjoint is the same operator with up <-> down and the flipped FIR (op/upfirdn2d.py:100-115)
        ctx.geom = (up, down, (kw - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                               kh - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1), input.shape, (out_h, out_w))
        return out.view(-1, c, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        grad_kernel, = ctx.saved_tensors
        (up_x, up_y), (down_x, down_y), g_pad, in_shape, (out_h, out_w) = ctx.geom
        g = upfirdn2d_op(grad_output.reshape(-1, out_h, out_w, 1), grad_kernel, down_x, down_y, up_x, up_y, *g_pad)
        return g.view(in_shape), None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """op/upfirdn2d.py:144-149: input [N, C, H, W], kernel [kh, kw]."""
    return _UpFirDn2dFn.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
