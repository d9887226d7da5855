"""Shared pieces of the frozen (inference-only) evaluation networks of the attribute-space traversal: eval-mode BatchNorm folded
into the convolution that feeds it, and one packed convolution = one tensor-core launch of libwgs_b200.  CUDA only."""
import torch

from . import conv as C


def fold_bn(weight, bias, bn):
    """conv -> eval BatchNorm as one conv: w' = w * gamma / sqrt(var + eps), b' = (b - mean) * gamma / sqrt(var + eps) + beta."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    b = bn.bias - bn.running_mean * s
    if bias is not None:
        b = b + bias * s
    return (weight * s.view(-1, 1, 1, 1)).float().contiguous(), b.float().contiguous()


def bn_affine(bn):
    """Eval BatchNorm as y = A x + B, each [1, C] (the operand-pack kernel applies it: a BatchNorm in FRONT of a zero-padded
    conv cannot be folded into the weights, the padding is applied after it)."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return s.float().view(1, -1).contiguous(), (bn.bias - bn.running_mean * s).float().view(1, -1).contiguous()


class PackedConv:
    """One frozen convolution: packed split32 weights + fp32 bias; call -> (fp32 NHWC output or None, split32 output or None)."""

    def __init__(self, weight, bias=None, stride=1, padding=0):
        if weight.device.type != 'cuda':
            raise RuntimeError('the evaluation networks run on CUDA only (no CPU fallback); call .cuda() first')
        self.co, self.ci, self.k, _ = weight.shape
        self.stride, self.padding = stride, padding
        self.w = C.pack_weights(weight.float().contiguous())
        self.b = bias.float().contiguous() if bias is not None else None

    def __call__(self, xs, act=0, out=None, accumulate=False, f32=True, split=False):
        n, h, w = xs.shape[0], xs.shape[1], xs.shape[2]
        oh = (h + 2 * self.padding - self.k) // self.stride + 1
        ow = (w + 2 * self.padding - self.k) // self.stride + 1
        out_split = torch.empty(n, oh, ow, C.chunks_of(self.co), 64, device=xs.device, dtype=torch.bfloat16) if split else None
        out = C.conv2d(xs, self.w, self.k, self.k, stride=self.stride, padding=self.padding, out=out,
                       no_f32=(out is None and not f32), cin=self.ci, beta=self.b, act=act, accumulate=accumulate,
                       out_split=out_split)
        return out, out_split


def need_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError('%s runs on CUDA tensors only (no CPU fallback); got %s' % (what, x.device))
