"""ResNet-18 feature extractor of the Reconstructor as ONE autograd node with a hand-scheduled forward / backward.

Every convolution (forward, data-gradient, weight-gradient) is a tensor-core launch; train-mode BatchNorm, the residual
add, ReLU and the split32 packing of the next conv's operand are fused into four bandwidth kernels (csrc/bn.cu), so an
activation is read / written a minimum number of times.  Semantics follow torchvision's resnet18 as used by
lib/reconstructor.py:54-79 (6-channel 7x7/2 stem, BasicBlocks, global average pool), BatchNorm in train mode
(batch statistics, running statistics updated with momentum 0.1) as lib/trainer.py:150 sets.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from . import conv as C
from . import wgrad as WG
from .reconstructor import conv_dgrad


def _bn_forward(y, bn, residual, relu, want_split, want_f32=True):
    """y [N,H,W,C] fp32 NHWC -> (z fp32 or None, zs split32 or None, (mean, rstd))."""
    n, h, w, c = y.shape
    R = n * h * w
    dev = y.device
    sums = torch.zeros(2, c, device=dev, dtype=torch.float32)
    _lib.call('wgs_bn_stats', _lib.ptr(y), R, c, _lib.ptr(sums[0]), _lib.ptr(sums[1]), _lib.stream())
    mean = torch.empty(c, device=dev, dtype=torch.float32)
    rstd = torch.empty(c, device=dev, dtype=torch.float32)
    _lib.call('wgs_bn_finalize', _lib.ptr(sums[0]), _lib.ptr(sums[1]), R, c, float(bn.eps), float(bn.momentum),
              _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.stream())
    bn.num_batches_tracked.add_(1)
    z = torch.empty_like(y) if want_f32 else None
    zs = torch.empty(n, h, w, C.chunks_of(c), 64, device=dev, dtype=torch.bfloat16) if want_split else None
    _lib.call('wgs_bn_act_fwd', _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(bn.weight.detach()),
              _lib.ptr(bn.bias.detach()), _lib.ptr(residual), int(relu), _lib.ptr(z), _lib.ptr(zs), R, c, _lib.stream())
    return z, zs, (mean, rstd)


def _bn_backward(dz, z, y, stats, gamma, relu, want_res):
    """-> (dys split32, dres fp32 or None, dgamma, dbeta)."""
    n, h, w, c = y.shape
    R = n * h * w
    dev = y.device
    mean, rstd = stats
    sums = torch.zeros(2, c, device=dev, dtype=torch.float32)
    _lib.call('wgs_bn_act_bwd_reduce', _lib.ptr(dz), _lib.ptr(z), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), int(relu),
              R, c, _lib.ptr(sums[0]), _lib.ptr(sums[1]), _lib.stream())
    dys = torch.empty(n, h, w, C.chunks_of(c), 64, device=dev, dtype=torch.bfloat16)
    dres = torch.empty_like(y) if want_res else None
    _lib.call('wgs_bn_act_bwd_apply', _lib.ptr(dz), _lib.ptr(z), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd),
              _lib.ptr(gamma.detach()), _lib.ptr(sums[0]), _lib.ptr(sums[1]), int(relu), R, c, _lib.ptr(dys), None,
              _lib.ptr(dres), _lib.stream())
    return dys, dres, sums[1], sums[0]


def _grad_target(w):
    """The trainer's flat, pre-zeroed gradient buffer when the wgrad kernel may accumulate straight into it.  Opt-in:
    only parameters re-homed by trainer.FlatParams carry `_wgs_flat_grad`; for everything else (a plain
    `loss.backward()` + torch.optim loop, torch.autograd.grad, hooks, GradScaler) the gradient is returned to autograd."""
    g = w.grad
    if not getattr(w, '_wgs_flat_grad', False):
        return None
    return g if (g is not None and g.is_contiguous() and g.dtype == torch.float32 and g.shape == w.shape) else None


def _conv(xs, w, stride, padding):
    co, ci, kh, kw = w.shape
    return C.conv2d(xs, C.pack_weights(w.detach()), kh, kw, stride=stride, padding=padding, cin=ci)


class ResNetFeatures(torch.autograd.Function):
    """features = avgpool(resnet18_trunk(x));  inputs: x (logical NCHW, channels-last), net (the _ResNet18 module),
    then every trainable tensor in `param_list(net)` order (so that autograd routes their gradients)."""

    @staticmethod
    def param_list(net):
        ps = [net.conv1.weight, net.bn1.weight, net.bn1.bias]
        for li in range(1, 5):
            for b in getattr(net, 'layer%d' % li):
                ps += [b.conv1.weight, b.bn1.weight, b.bn1.bias, b.conv2.weight, b.bn2.weight, b.bn2.bias]
                if hasattr(b, 'downsample'):
                    ps += [b.downsample[0].weight, b.downsample[1].weight, b.downsample[1].bias]
        return ps

    @staticmethod
    def forward(ctx, x, net, *params):
        if not x.is_cuda:
            raise RuntimeError('Reconstructor runs on CUDA tensors only (no CPU fallback); got %s' % x.device)
        tape = {}
        n, ci, h, w = x.shape
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        # stem: im2col -> 1-tap GEMM
        w1 = net.conv1.weight
        co, _, kh, kw = w1.shape
        oh, ow = (h + 6 - kh) // 2 + 1, (w + 6 - kw) // 2 + 1
        chunks = (kh * kw * ci + 31) // 32
        xcol = torch.empty(n, oh, ow, chunks, 64, device=x.device, dtype=torch.bfloat16)
        _lib.call('wgs_im2col_split32', _lib.ptr(x_nhwc), n, h, w, ci, kh, kw, 2, 3, oh, ow, _lib.ptr(xcol), _lib.stream())
        wmat = w1.detach().permute(0, 2, 3, 1).reshape(co, kh * kw * ci, 1, 1)
        y0 = C.conv2d(xcol, C.pack_weights(wmat), 1, 1, cin=kh * kw * ci)
        z0, _, st0 = _bn_forward(y0, net.bn1, None, True, want_split=False)
        ph, pw = (oh + 1) // 2, (ow + 1) // 2
        cur = torch.empty(n, ph, pw, co, device=x.device, dtype=torch.float32)             # NHWC fp32
        pool_idx = torch.empty(n, ph, pw, co, device=x.device, dtype=torch.uint8)
        cur_s = torch.empty(n, ph, pw, C.chunks_of(co), 64, device=x.device, dtype=torch.bfloat16)
        _lib.call('wgs_maxpool3s2_fwd', _lib.ptr(z0), n, oh, ow, co, _lib.ptr(cur), _lib.ptr(pool_idx), _lib.ptr(cur_s),
                  _lib.stream())
        tape['stem'] = (xcol, y0, z0, st0, pool_idx, (n, ci, h, w))
        tape['blocks'] = []
        for li in range(1, 5):
            for b in getattr(net, 'layer%d' % li):
                s = b.stride
                y1 = _conv(cur_s, b.conv1.weight, s, 1)
                z1, z1s, st1 = _bn_forward(y1, b.bn1, None, True, want_split=True)
                y2 = _conv(z1s, b.conv2.weight, 1, 1)
                if hasattr(b, 'downsample'):
                    yd = _conv(cur_s, b.downsample[0].weight, s, 0)
                    idt, _, std = _bn_forward(yd, b.downsample[1], None, False, want_split=False)
                else:
                    yd, std, idt = None, None, cur
                out, outs, st2 = _bn_forward(y2, b.bn2, idt, True, want_split=True)
                tape['blocks'].append((b, cur_s, cur.shape, y1, z1, z1s, st1, y2, out, st2, yd, std))
                cur, cur_s = out, outs
        feat = cur.mean(dim=(1, 2))
        tape['final_shape'] = cur.shape
        ctx.tape, ctx.net = tape, net
        ctx.need_dx = ctx.needs_input_grad[0]
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        tape, net = ctx.tape, ctx.net
        grads = {}
        n, fh, fw, fc = tape['final_shape']
        dcur = (dfeat.contiguous().view(n, 1, 1, fc) / float(fh * fw)).expand(n, fh, fw, fc).contiguous()
        for (b, xs, x_shape, y1, z1, z1s, st1, y2, out, st2, yd, std) in reversed(tape['blocks']):
            s = b.stride
            _, xh, xw, xc = x_shape
            dy2s, dres, dg2, db2 = _bn_backward(dcur, out, y2, st2, b.bn2.weight, True, want_res=True)
            grads[b.bn2.weight], grads[b.bn2.bias] = dg2, db2
            w2 = b.conv2.weight
            grads[w2] = WG.conv_wgrad(z1s, dy2s, tuple(w2.shape), 1, 1, out=_grad_target(w2))
            dz1 = conv_dgrad(dy2s, w2, (y1.shape[1], y1.shape[2]), 1, 1)
            dy1s, _, dg1, db1 = _bn_backward(dz1, z1, y1, st1, b.bn1.weight, True, want_res=False)
            grads[b.bn1.weight], grads[b.bn1.bias] = dg1, db1
            w1 = b.conv1.weight
            grads[w1] = WG.conv_wgrad(xs, dy1s, tuple(w1.shape), s, 1, out=_grad_target(w1))
            if yd is not None:
                dyds, _, dgd, dbd = _bn_backward(dres, None, yd, std, b.downsample[1].weight, False, want_res=False)
                grads[b.downsample[1].weight], grads[b.downsample[1].bias] = dgd, dbd
                wd = b.downsample[0].weight
                grads[wd] = WG.conv_wgrad(xs, dyds, tuple(wd.shape), s, 0, out=_grad_target(wd))
                dx = conv_dgrad(dy1s, w1, (xh, xw), s, 1)
                dx = conv_dgrad(dyds, wd, (xh, xw), s, 0, out=dx, accumulate=True)
            else:
                dx = conv_dgrad(dy1s, w1, (xh, xw), s, 1, out=dres, accumulate=True)     # dx = dres + conv^T(dy1)
            dcur = dx
        xcol, y0, z0, st0, pool_idx, (n, ci, h, w) = tape['stem']
        dz0 = torch.empty_like(z0)
        dcur = dcur.contiguous()           # (held in a name: _lib.ptr() of a temporary would free it before the launch)
        _lib.call('wgs_maxpool3s2_bwd', _lib.ptr(dcur), _lib.ptr(pool_idx), z0.shape[0], z0.shape[1], z0.shape[2],
                  z0.shape[3], _lib.ptr(dz0), _lib.stream())
        dy0s, _, dg0, db0 = _bn_backward(dz0, z0, y0, st0, net.bn1.weight, True, want_res=False)
        grads[net.bn1.weight], grads[net.bn1.bias] = dg0, db0
        w1 = net.conv1.weight
        co, _, kh, kw = w1.shape
        dwm = WG.conv_wgrad(xcol, dy0s, (co, kh * kw * ci, 1, 1), 1, 0)
        grads[w1] = dwm.reshape(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
        dx = conv_dgrad(dy0s, w1, (h, w), 2, 3).permute(0, 3, 1, 2) if ctx.need_dx else None
        ctx.tape = None
        plist = ResNetFeatures.param_list(net)
        return (dx, None) + tuple(grads.get(p) for p in plist)
