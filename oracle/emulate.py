"""Oracle with the ARITHMETIC MODEL of the libwgs_b200 convolutions (test infrastructure, CPU torch).

The tensor-core kernels take fp32 operands split into bf16 hi + lo (x = hi + lo to 2^-17) and accumulate in fp32.  This
module re-runs the oracle with exactly that operand rounding applied in front of every dense convolution of the forward
pass (straight-through in the backward pass, which stays exact fp32).  It answers one question for the gradient parity
tests: how far do the whole-graph gradients of the paired step move when the forward activations move by the kernels'
rounding (~1e-5 relative)?  Measured (tools/grad_probe.py, profiles/r02_gradient_conditioning.md): ~1e-2, i.e. the graph
amplifies forward perturbations ~1000x into its gradients - train-mode BatchNorm removes the mean / scale component of
every gradient, so what is left is a small residual of large cancelling terms.  The reference on a GPU (TF32 cuDNN
convolutions, forward error ~8e-4) sits two orders of magnitude further out still.
"""
import contextlib
import types

import torch
import torch.nn.functional as F


class _Split17(torch.autograd.Function):
    """x -> bf16(x) + bf16(x - bf16(x)); identity gradient."""

    @staticmethod
    def forward(ctx, x):
        if x.dtype != torch.float32:
            return x
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _namespace():
    ns = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith('__')})

    def conv2d(x, w, *a, **k):
        groups = k.get('groups', a[4] if len(a) > 4 else 1)
        if groups == x.shape[1] and w.shape[1] == 1:          # depthwise FIR (upfirdn2d): CUDA-core fp32 in the kernels
            return F.conv2d(x, w, *a, **k)
        return F.conv2d(_Split17.apply(x), _Split17.apply(w), *a, **k)

    def conv_transpose2d(x, w, *a, **k):
        return F.conv_transpose2d(_Split17.apply(x), _Split17.apply(w), *a, **k)

    ns.conv2d, ns.conv_transpose2d = conv2d, conv_transpose2d
    return ns


@contextlib.contextmanager
def split17_convs(*modules):
    """Inside the block, the given oracle modules (e.g. oracle.stylegan2, oracle.reconstructor) see F.conv2d /
    F.conv_transpose2d with bf16 hi+lo operand rounding."""
    saved = [(m, m.F) for m in modules]
    ns = _namespace()
    try:
        for m in modules:
            m.F = ns
        yield
    finally:
        for m, f in saved:
            m.F = f
