"""ResNet-18 feature extractor of the Reconstructor as ONE autograd node with a hand-scheduled forward / backward.

Every convolution (forward, data-gradient, weight-gradient) is a tensor-core launch; train-mode BatchNorm, the residual
add, ReLU and the split32 packing of the next conv's operand are fused into three bandwidth kernels per conv
(csrc/bn.cu: shifted statistics, apply with the finalize folded in; backward reduce + apply), so an activation is read /
written a minimum number of times.  Semantics follow torchvision's resnet18 as used by lib/reconstructor.py:54-79
(6-channel 7x7/2 stem, BasicBlocks, global average pool), BatchNorm in train mode (batch statistics, running statistics
updated with momentum 0.1) as lib/trainer.py:150 sets.

Launch diet (the weights change every step, so none of this can be cached across steps):
  * all 41 weight packs of a step (20 forward, 20 data-gradient incl. the phase-merged strided ones, the stem's im2col
    matrix) are ONE grouped launch from the torch layout (wgs_pack_weights_group) instead of a permute-copy + a pack each;
  * the per-BatchNorm reduction buffers are slices of one zero-filled pool per pass; when the parameters live in the
    trainer's flat buffers (trainer.FlatParams) the backward reductions ARE the d gamma / d beta gradient views, so there
    is neither a zero fill nor an accumulate per parameter;
  * num_batches_tracked is bumped by one multi-tensor add.
"""
import os

import torch

from . import _lib
from . import conv as C
from . import wgrad as WG
from .reconstructor import conv_dgrad


def _flat(p):
    """The parameter's flat-buffer gradient view (pre-zeroed by the trainer) when kernels may accumulate straight into it.
    Opt-in: only parameters re-homed by trainer.FlatParams carry `_wgs_flat_grad`; for everything else (a plain
    `loss.backward()` + torch.optim loop, torch.autograd.grad, hooks, GradScaler) gradients are returned to autograd."""
    g = p.grad
    if not getattr(p, '_wgs_flat_grad', False):
        return None
    return g if (g is not None and g.is_contiguous() and g.dtype == torch.float32 and g.shape == p.shape) else None


class _Pool:
    """Slices of one zero-filled fp32 buffer (one fill launch instead of one per BatchNorm)."""

    def __init__(self, n, device):
        self.buf = torch.zeros(n, device=device, dtype=torch.float32)
        self.off = 0

    def take(self, n):
        out = self.buf[self.off: self.off + n]
        self.off += (n + 3) // 4 * 4
        return out


def _bn_list(net):
    bns = [net.bn1]
    for li in range(1, 5):
        for b in getattr(net, 'layer%d' % li):
            bns += [b.bn1, b.bn2] + ([b.downsample[1]] if hasattr(b, 'downsample') else [])
    return bns


# 1 = BatchNorm statistics formed in the conv epilogue (wgs_conv_desc.stat_sum) instead of by a separate pass.  Measured on one
# B200 (profiles/r02_launch_list_summary.md): the pass it removes costs 0.32 ms per step, the epilogue work it adds 0.33 ms
# (the multi-tile kernels of the stem / layer 1 are epilogue-bound: 73 -> 100 us per launch) - a wash, so the separate pass
# stays the default and this is an A/B switch.
# The projection shortcut of layer 2 / 3 / 4's first block on a forked stream beside the main branch, for blocks of at most this
# many output elements (latency-bound maps: the 128 x 128 inputs of config 4 gain 2.5 %, 6.96 / 7.06 -> 6.82 ms; on 1024 x 1024
# inputs every launch fills the GPU and the fork costs 0.4 %, 15.94 / 15.98 -> 16.03 / 16.04 ms); 0 = never
SHORTCUT_FORK_MAX = int(os.environ.get('WGS_SHORTCUT_FORK_MAX', '1000000'))
EPILOGUE_STATS = os.environ.get('WGS_EPILOGUE_STATS', '0') == '1'


def _stat_slots(bn, pool, shifts, out_hw):
    """(sum, sumsq, shift) for a conv whose epilogue forms the BatchNorm statistics of its output, or None when the output
    map is too small for that path (< 128 pixels per image: several images share a tile)."""
    if not EPILOGUE_STATS or out_hw[0] * out_hw[1] < 128 or bn.num_features % 16 != 0:
        return None
    return pool.take(bn.num_features), pool.take(bn.num_features), shifts[id(bn)]


def _bn_forward(y, bn, residual, relu, want_split, pool, want_f32=True, pre=None):
    """y [N,H,W,C] fp32 NHWC -> (z fp32 or None, zs split32 or None, (mean, rstd)).  pre = (sum, sumsq, shift) when the conv
    that produced y already accumulated the statistics in its epilogue."""
    n, h, w, c = y.shape
    R = n * h * w
    dev = y.device
    if pre is not None:
        s0, s1, shift = pre
    else:
        s0, s1, shift = pool.take(c), pool.take(c), None
        _lib.call('wgs_bn_stats', _lib.ptr(y), R, c, _lib.ptr(s0), _lib.ptr(s1), _lib.stream())
    stats = torch.empty(2, c, device=dev, dtype=torch.float32)
    z = torch.empty_like(y) if want_f32 else None
    zs = torch.empty(n, h, w, C.chunks_of(c), 64, device=dev, dtype=torch.bfloat16) if want_split else None
    gamma, beta = bn.weight.detach(), bn.bias.detach()
    _lib.call('wgs_bn_fwd_fused', _lib.ptr(y), _lib.ptr(s0), _lib.ptr(s1), _lib.ptr(shift), R, c, float(bn.eps), float(bn.momentum),
              _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(residual), int(relu), _lib.ptr(z), _lib.ptr(zs),
              _lib.ptr(stats[0]), _lib.ptr(stats[1]), _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.stream())
    return z, zs, (stats[0], stats[1])


def _bn_backward(dz, z, y, stats, bn, relu, want_res, pool, grads):
    """-> (dys split32, dres fp32 or None); d gamma / d beta land in the flat gradient views or in `grads`."""
    n, h, w, c = y.shape
    R = n * h * w
    dev = y.device
    mean, rstd = stats
    d_beta, d_gamma = _flat(bn.bias), _flat(bn.weight)
    if d_beta is None or d_gamma is None:
        d_beta, d_gamma = pool.take(c), pool.take(c)
        grads[bn.weight], grads[bn.bias] = d_gamma, d_beta
    gamma, beta = bn.weight.detach(), bn.bias.detach()
    _lib.call('wgs_bn_act_bwd_reduce', _lib.ptr(dz), _lib.ptr(z), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
              _lib.ptr(beta), int(relu), R, c, _lib.ptr(d_beta), _lib.ptr(d_gamma), _lib.stream())
    dys = torch.empty(n, h, w, C.chunks_of(c), 64, device=dev, dtype=torch.bfloat16)
    dres = torch.empty_like(y) if want_res else None
    _lib.call('wgs_bn_act_bwd_apply', _lib.ptr(dz), _lib.ptr(z), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd),
              _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(d_beta), _lib.ptr(d_gamma), int(relu), R, c, _lib.ptr(dys), None,
              _lib.ptr(dres), _lib.stream())
    return dys, dres


def _pack_all(net, stem_sub=None):
    """One grouped launch: {(weight, 'fwd' | 'bwd'): packed}.  stem_sub = (offset, count): the stem's data-gradient pack
    covers that slice of its input channels only (the shifted image of a pair)."""
    specs, keys = [], []

    def add(w, stride, padding, stem=False):
        wd = w.detach()
        specs.append((wd, C.PACK_S2D, padding) if stem else (wd, C.PACK_FWD, None))
        keys.append((id(w), 'fwd'))
        extra = (stride, padding) + (tuple(stem_sub) if (stem and stem_sub) else ())
        specs.append((wd, C.PACK_MERGED_DGRAD, extra) if stride > 1 else (wd, C.PACK_TRANSPOSED, None))
        keys.append((id(w), 'bwd'))

    add(net.conv1.weight, 2, 3, stem=True)
    for li in range(1, 5):
        for b in getattr(net, 'layer%d' % li):
            add(b.conv1.weight, b.stride, 1)
            add(b.conv2.weight, 1, 1)
            if hasattr(b, 'downsample'):
                add(b.downsample[0].weight, b.stride, 0)
    return dict(zip(keys, C.pack_weights_group(specs)))


def prepack(net, stem_sub, stream):
    """Form every weight operand of the coming step on `stream` (a side stream forked from the current one) so that the
    grouped pack - 0.17 ms that depends on nothing but the weights - runs underneath the RBF kernel / mapping network /
    low-resolution generator layers, which leave most of the GPU idle.  ResNetFeatures.forward picks the result up (and
    waits for it) when the slice matches."""
    cur = torch.cuda.current_stream()
    stream.wait_stream(cur)                      # the previous step's Adam update and last readers of the old packs are done
    with torch.cuda.stream(stream):
        packs = _pack_all(net, stem_sub)
        done = torch.cuda.Event()
        done.record(stream)
    net._wgs_prepacked = (stem_sub, packs, done)


_S2D_INDEX = {}


def _s2d_weight_grad(dws, ci, k, S, G):
    """dws [co, 4*ci, S, S] (channel (py*2+px)*ci + c, tap (ty, tx) = kernel element (2*ty+py-G, 2*tx+px-G)) -> [co, ci, k, k]."""
    key = (ci, k, S, G, str(dws.device))
    idx = _S2D_INDEX.get(key)
    if idx is None:
        flat = []
        for c in range(ci):
            for ky in range(k):
                for kx in range(k):
                    py, ty = (ky + G) % 2, (ky + G) // 2
                    px, tx = (kx + G) % 2, (kx + G) // 2
                    flat.append((((py * 2 + px) * ci + c) * S + ty) * S + tx)
        idx = _S2D_INDEX[key] = torch.tensor(flat, device=dws.device)
    co = dws.shape[0]
    return dws.reshape(co, -1).index_select(1, idx).reshape(co, ci, k, k)


class ResNetFeatures(torch.autograd.Function):
    """features = avgpool(resnet18_trunk(cat(x, x2)));  inputs: x, x2 (logical NCHW, channels-last; x2 may be None), net
    (the _ResNet18 module), then every trainable tensor in `param_list(net)` order (so that autograd routes their
    gradients).  With two inputs the channel concatenation of lib/reconstructor.py:72 is folded into the stem's operand
    pack, and when only x2 needs a gradient (the paired step: x1 = G(z) is detached) the stem's data gradient is formed
    for x2's channels alone and returned as is - no cat, no slice copies."""

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
    def forward(ctx, x, x2, net, *params):
        if not x.is_cuda or (x2 is not None and not x2.is_cuda):
            raise RuntimeError('Reconstructor runs on CUDA tensors only (no CPU fallback); got %s' % x.device)
        tape = {}
        n, c1, h, w = x.shape
        c2 = x2.shape[1] if x2 is not None else 0
        ci = c1 + c2
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        x2_nhwc = x2.permute(0, 2, 3, 1).contiguous() if x2 is not None else None
        ctx.need_dx = (ctx.needs_input_grad[0], x2 is not None and ctx.needs_input_grad[1])
        ctx.split = (c1, c2)
        stem_sub = (c1, c2) if ctx.need_dx == (False, True) else None
        pre = getattr(net, '_wgs_prepacked', None)
        net._wgs_prepacked = None
        if pre is not None and pre[0] == stem_sub:
            packs = pre[1]
            torch.cuda.current_stream().wait_event(pre[2])
        else:
            packs = _pack_all(net, stem_sub)
        bns = _bn_list(net)
        pool = _Pool(sum(2 * ((b.num_features + 3) // 4 * 4) for b in bns), x.device)

        # shift vectors of the epilogue statistics: a snapshot of every running mean (|mean - shift| ~ std after the first
        # steps, so E[(y-s)^2] - E[y-s]^2 keeps its digits; the snapshot is one cat, taken before any BatchNorm updates them)
        shifts, off = {}, 0
        if EPILOGUE_STATS:
            snap = torch.cat([b.running_mean for b in bns])
            for b in bns:
                shifts[id(b)] = snap[off: off + b.num_features]
                off += b.num_features

        def conv(xs, wt, stride, padding, bn=None):
            co, cin, kh, kw = wt.shape
            oh_, ow_ = (xs.shape[1] + 2 * padding - kh) // stride + 1, (xs.shape[2] + 2 * padding - kw) // stride + 1
            pre = _stat_slots(bn, pool, shifts, (oh_, ow_)) if bn is not None else None
            y = C.conv2d(xs, packs[(id(wt), 'fwd')], kh, kw, stride=stride, padding=padding, cin=cin, split_k=True, stats=pre)
            return y, pre

        # stem 7x7 / 2 on 6 channels: 2x2 space-to-depth (24 channels, one split32 chunk) -> stride-1 4x4-tap conv on the
        # multi-tile halo kernel (taps resident in shared memory); no im2col (it wrote 1280 B per output pixel)
        w1 = net.conv1.weight
        co, _, kh, kw = w1.shape
        oh, ow = (h + 6 - kh) // 2 + 1, (w + 6 - kw) // 2 + 1
        xs2d = C.s2d_pack_split32(x_nhwc, x2_nhwc)
        taps, S = C.s2d_taps(kh, 3)
        y0 = torch.empty(n, oh, ow, co, device=x.device, dtype=torch.float32)
        bn1 = net.bn1
        pre0 = _stat_slots(bn1, pool, shifts, (oh, ow))
        C.conv_taps(xs2d, packs[(id(w1), 'fwd')], taps, y0, grid=(oh, ow), cin=kh * kw * ci,
                    algo_macs_per_pixel=kh * kw * ci * co, stats=pre0)
        # bn1 + relu + maxpool in one pass over y0: the normalised 512^2 activation is never stored
        if pre0 is not None:
            s0, s1, sh0 = pre0
        else:
            s0, s1, sh0 = pool.take(co), pool.take(co), None
            _lib.call('wgs_bn_stats', _lib.ptr(y0), n * oh * ow, co, _lib.ptr(s0), _lib.ptr(s1), _lib.stream())
        st0 = torch.empty(2, co, device=x.device, dtype=torch.float32)
        ph, pw = (oh + 1) // 2, (ow + 1) // 2
        cur = torch.empty(n, ph, pw, co, device=x.device, dtype=torch.float32)             # NHWC fp32
        pool_idx = torch.empty(n, ph, pw, co, device=x.device, dtype=torch.uint8)
        cur_s = torch.empty(n, ph, pw, C.chunks_of(co), 64, device=x.device, dtype=torch.bfloat16)
        _lib.call('wgs_bn_pool_fwd', _lib.ptr(y0), _lib.ptr(s0), _lib.ptr(s1), _lib.ptr(sh0), n, oh, ow, co, float(bn1.eps),
                  float(bn1.momentum), _lib.ptr(bn1.weight.detach()), _lib.ptr(bn1.bias.detach()), _lib.ptr(cur),
                  _lib.ptr(pool_idx), _lib.ptr(cur_s), _lib.ptr(st0[0]), _lib.ptr(st0[1]), _lib.ptr(bn1.running_mean),
                  _lib.ptr(bn1.running_var), _lib.stream())
        tape['stem'] = (xs2d, y0, st0, pool_idx, (n, ci, h, w))
        tape['blocks'] = []
        main = torch.cuda.current_stream()
        aux_stream = C._phase_streams(x.device, 4)[3] if (SHORTCUT_FORK_MAX > 0 and C.PROFILE is None) else None
        for li in range(1, 5):
            for b in getattr(net, 'layer%d' % li):
                s = b.stride
                small = cur.shape[0] * (cur.shape[1] // s) * (cur.shape[2] // s) * b.conv2.weight.shape[0] <= SHORTCUT_FORK_MAX
                aux = aux_stream if small else None
                if hasattr(b, 'downsample'):
                    # the projection shortcut (1x1 / 2 conv + BatchNorm: three launches that depend on the block input alone)
                    # runs on a forked stream beside conv1 -> bn1 -> conv2 and is joined in front of bn2
                    if aux is not None:
                        fork = torch.cuda.Event()
                        fork.record(main)
                        aux.wait_event(fork)
                    with torch.cuda.stream(aux if aux is not None else main):
                        yd, pd = conv(cur_s, b.downsample[0].weight, s, 0, b.downsample[1])
                        idt, _, std = _bn_forward(yd, b.downsample[1], None, False, False, pool, pre=pd)
                        if aux is not None:
                            joined = torch.cuda.Event()
                            joined.record(aux)
                else:
                    yd, std, idt = None, None, cur
                y1, p1 = conv(cur_s, b.conv1.weight, s, 1, b.bn1)
                # (the mid-block activation is only needed as the next conv's operand: its ReLU mask is re-derived from y1 in
                # the backward pass, so the fp32 copy is neither written here nor read there)
                z1, z1s, st1 = _bn_forward(y1, b.bn1, None, True, True, pool, want_f32=False, pre=p1)
                y2, p2 = conv(z1s, b.conv2.weight, 1, 1, b.bn2)
                if yd is not None and aux is not None:
                    main.wait_event(joined)
                out, outs, st2 = _bn_forward(y2, b.bn2, idt, True, True, pool, pre=p2)
                tape['blocks'].append((b, cur_s, cur.shape, y1, z1, z1s, st1, y2, out, st2, yd, std))
                cur, cur_s = out, outs
        torch._foreach_add_([b.num_batches_tracked for b in bns], 1)
        feat = cur.mean(dim=(1, 2))
        tape['final_shape'] = cur.shape
        tape['packs'] = packs
        ctx.tape, ctx.net = tape, net
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        tape, net = ctx.tape, ctx.net
        packs = tape['packs']
        grads = {}
        bns = _bn_list(net)
        pool = _Pool(sum(2 * ((b.num_features + 3) // 4 * 4) for b in bns), dfeat.device)
        n, fh, fw, fc = tape['final_shape']
        dcur = (dfeat.contiguous().view(n, 1, 1, fc) / float(fh * fw)).expand(n, fh, fw, fc).contiguous()
        main = torch.cuda.current_stream()
        aux_stream = C._phase_streams(dfeat.device, 4)[3] if (SHORTCUT_FORK_MAX > 0 and C.PROFILE is None) else None

        # Weight gradients hang off the data-gradient chain: nothing on the way back to the generator waits for them.
        # With a side stream from the trainer (net._wgs_wgrad_stream, low priority) they are launched there as soon as
        # their operands exist and fill whatever the main chain leaves idle; operands are kept alive in net._wgs_keep until
        # the trainer joins the stream (the caching allocator knows nothing about this second reader).
        side = getattr(net, '_wgs_wgrad_stream', None)
        keep = []
        if side is not None:
            net._wgs_keep = keep

        def wgrad(xs, dys, wt, stride, padding):
            out = _flat(wt)
            if side is None or out is None:
                grads[wt] = WG.conv_wgrad(xs, dys, tuple(wt.shape), stride, padding, out=out)
                return
            ready = torch.cuda.Event()
            ready.record()
            side.wait_event(ready)
            with torch.cuda.stream(side):
                WG.conv_wgrad(xs, dys, tuple(wt.shape), stride, padding, out=out)
            keep.append((xs, dys))

        def dgrad(dys, wt, in_hw, stride, padding, out=None, accumulate=False):
            return conv_dgrad(dys, wt, in_hw, stride, padding, out=out, accumulate=accumulate, packed=packs[(id(wt), 'bwd')],
                              split_k=True)

        for (b, xs, x_shape, y1, z1, z1s, st1, y2, out, st2, yd, std) in reversed(tape['blocks']):
            s = b.stride
            _, xh, xw, xc = x_shape
            dy2s, dres = _bn_backward(dcur, out, y2, st2, b.bn2, True, True, pool, grads)
            aux = aux_stream if y2.numel() <= SHORTCUT_FORK_MAX else None
            if yd is not None and aux is not None:
                # shortcut branch on the forked stream: BatchNorm backward, weight gradient, data gradient into dx; the main
                # branch accumulates conv1's data gradient on top of it after the join
                fork = torch.cuda.Event()
                fork.record(main)
                aux.wait_event(fork)
                with torch.cuda.stream(aux):
                    dyds, _ = _bn_backward(dres, None, yd, std, b.downsample[1], False, False, pool, grads)
                    wgrad(xs, dyds, b.downsample[0].weight, s, 0)
                    dx = dgrad(dyds, b.downsample[0].weight, (xh, xw), s, 0)
                    joined = torch.cuda.Event()
                    joined.record(aux)
            wgrad(z1s, dy2s, b.conv2.weight, 1, 1)
            dz1 = dgrad(dy2s, b.conv2.weight, (y1.shape[1], y1.shape[2]), 1, 1)
            dy1s, _ = _bn_backward(dz1, z1, y1, st1, b.bn1, True, False, pool, grads)
            wgrad(xs, dy1s, b.conv1.weight, s, 1)
            if yd is not None and aux is not None:
                main.wait_event(joined)
                dx = dgrad(dy1s, b.conv1.weight, (xh, xw), s, 1, out=dx, accumulate=True)
            elif yd is not None:
                dyds, _ = _bn_backward(dres, None, yd, std, b.downsample[1], False, False, pool, grads)
                wgrad(xs, dyds, b.downsample[0].weight, s, 0)
                dx = dgrad(dy1s, b.conv1.weight, (xh, xw), s, 1)
                dx = dgrad(dyds, b.downsample[0].weight, (xh, xw), s, 0, out=dx, accumulate=True)
            else:
                dx = dgrad(dy1s, b.conv1.weight, (xh, xw), s, 1, out=dres, accumulate=True)     # dx = dres + conv^T(dy1)
            dcur = dx
        xs2d, y0, st0, pool_idx, (n, ci, h, w) = tape['stem']
        dcur = dcur.contiguous()           # (held in a name: _lib.ptr() of a temporary would free it before the launch)
        bn1 = net.bn1
        _, oh, ow, c0 = y0.shape
        d_beta, d_gamma = _flat(bn1.bias), _flat(bn1.weight)
        if d_beta is None or d_gamma is None:
            d_beta, d_gamma = pool.take(c0), pool.take(c0)
            grads[bn1.weight], grads[bn1.bias] = d_gamma, d_beta
        g1, b1 = bn1.weight.detach(), bn1.bias.detach()
        _lib.call('wgs_bn_pool_bwd_reduce', _lib.ptr(dcur), _lib.ptr(pool_idx), _lib.ptr(y0), _lib.ptr(st0[0]), _lib.ptr(st0[1]),
                  _lib.ptr(g1), _lib.ptr(b1), n, oh, ow, c0, _lib.ptr(d_beta), _lib.ptr(d_gamma), _lib.stream())
        dy0s = torch.empty(n, oh, ow, C.chunks_of(c0), 64, device=y0.device, dtype=torch.bfloat16)
        _lib.call('wgs_bn_pool_bwd_apply', _lib.ptr(dcur), _lib.ptr(pool_idx), _lib.ptr(y0), _lib.ptr(st0[0]), _lib.ptr(st0[1]),
                  _lib.ptr(g1), _lib.ptr(b1), _lib.ptr(d_beta), _lib.ptr(d_gamma), n, oh, ow, c0, _lib.ptr(dy0s), _lib.stream())
        w1 = net.conv1.weight
        co, _, kh, kw = w1.shape
        # weight gradient in the space-to-depth form [co, 4*ci, S, S], gathered back to [co, ci, kh, kw]
        S, a_min, G = C.s2d_geometry(kh, 3)
        flat1 = _flat(w1)
        if side is not None and flat1 is not None:
            ready = torch.cuda.Event()
            ready.record()
            side.wait_event(ready)
            with torch.cuda.stream(side):
                dws = WG.conv_wgrad(xs2d, dy0s, (co, 4 * ci, S, S), 1, -a_min)
                flat1.add_(_s2d_weight_grad(dws, ci, kh, S, G))
            keep.append((xs2d, dy0s))
        else:
            dws = WG.conv_wgrad(xs2d, dy0s, (co, 4 * ci, S, S), 1, -a_min)
            grads[w1] = _s2d_weight_grad(dws, ci, kh, S, G)
        dx1 = dx2 = None
        c1, c2 = ctx.split
        if ctx.need_dx == (False, True):
            dx2 = C.conv_dgrad_merged(dy0s, w1, (h, w), 2, 3, w_merged=packs[(id(w1), 'bwd')], ci_sub=(c1, c2)).permute(0, 3, 1, 2)
        elif ctx.need_dx[0] or ctx.need_dx[1]:
            dx = dgrad(dy0s, w1, (h, w), 2, 3).permute(0, 3, 1, 2)
            dx1 = dx[:, :c1] if ctx.need_dx[0] else None
            dx2 = dx[:, c1:] if ctx.need_dx[1] else None
        ctx.tape = None
        plist = ResNetFeatures.param_list(net)
        return (dx1, dx2, None) + tuple(grads.get(p) for p in plist)
