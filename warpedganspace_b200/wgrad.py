"""Weight gradient of the Reconstructor convolutions.

dW[co,ci,ky,kx] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s - p + ky, ox*s - p + kx, ci]
The contraction runs over pixels, so both operands are "MN-major" for the tensor core; see csrc/wgrad.cu.
"""
import torch

from . import _lib
from . import conv as C


def conv_wgrad(xs, dys, w_shape, stride, padding):
    """xs: split32 input [N,H,W,ci_chunks,64]; dys: split32 output gradient [N,OH,OW,co_chunks,64].
    Returns fp32 [Co, Ci, kh, kw]."""
    co, ci, kh, kw = w_shape
    n, h, wd, ci_chunks, _ = xs.shape
    _, oh, ow, co_chunks, _ = dys.shape
    dw = torch.zeros(kh * kw, co_chunks * 32, ci_chunks * 32, device=xs.device, dtype=torch.float32)
    e0 = C._prof_begin()
    _lib.call('wgs_conv_wgrad_split32', _lib.ptr(xs), n, h, wd, ci_chunks, _lib.ptr(dys), oh, ow, co_chunks,
              kh, kw, stride, padding, _lib.ptr(dw), _lib.stream())
    C._prof_end(e0, 'wgrad', 2.0 * n * oh * ow * co * ci * kh * kw)
    return dw[:, :co, :ci].reshape(kh, kw, co, ci).permute(2, 3, 0, 1).contiguous()
