"""Weight gradient of the Reconstructor convolutions.

dW[co,ci,ky,kx] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s - p + ky, ox*s - p + kx, ci]
The contraction runs over pixels, so both operands are "MN-major" for the tensor core; see csrc/wgrad.cu.
"""
import torch

from . import _lib
from . import conv as C


def conv_wgrad(xs, dys, w_shape, stride, padding, out=None):
    """xs: split32 input [N,H,W,ci_chunks,64]; dys: split32 output gradient [N,OH,OW,co_chunks,64].
    Returns fp32 [Co, Ci, kh, kw]; with `out` (a zero-initialised or accumulating contiguous [Co,Ci,kh,kw] tensor,
    e.g. the parameter's flat-buffer .grad) the kernel accumulates straight into it and None is returned."""
    co, ci, kh, kw = w_shape
    n, h, wd, ci_chunks, _ = xs.shape
    _, oh, ow, co_chunks, _ = dys.shape
    dw = out if out is not None else torch.zeros(co, ci, kh, kw, device=xs.device, dtype=torch.float32)
    assert dw.is_contiguous() and dw.dtype == torch.float32 and tuple(dw.shape) == (co, ci, kh, kw)
    e0 = C._prof_begin()
    _lib.call('wgs_conv_wgrad_split32', _lib.ptr(xs), n, h, wd, ci_chunks, _lib.ptr(dys), oh, ow, co_chunks,
              kh, kw, stride, padding, _lib.ptr(dw), 1, co, ci, _lib.stream())
    C._prof_end(e0, 'wgrad', 2.0 * n * oh * ow * co * ci * kh * kw)
    return None if out is not None else dw
