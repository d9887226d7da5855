"""ProgGAN, SNGAN and BigGAN generators with every convolution on the tcgen05 tap-list kernel.

State-dict keys equal the reference modules' (models/ProgGAN/model.py:65-95, models/SNGAN/sn_gen_resnet.py:81-112,
models/BigGAN/BigGAN.py:54-243), so released checkpoints load with ``load_state_dict``.  The generators are
frozen during WarpedGANSpace training: weights are folded once per ``plan()`` (ProgGAN WScale into the conv
weights; BigGAN spectral norm W/sigma — the reference redoes a non-updating power iteration on every forward,
models/BigGAN/layers.py:84-96; eval-mode BatchNorm into an affine map) and only the data gradient is propagated.
ProgGAN and BigGAN are hand-scheduled kernel chains: every convolution is a tensor-core launch with bias / activation in
its epilogue, the glue between convolutions (pixel norm, class-conditional BatchNorm + ReLU, operand packing) is ONE
element-wise pass per conv (csrc/proggan.cu, csrc/ccbn.cu), and nearest x2 + conv3x3 runs as four output-phase 2x2 convs
over the low-resolution operand (2.25x fewer MACs).  Still library calls: BigGAN's attention matrix products / soft-max
(2 % of its FLOPs) and the tiny linears; SNGAN (config 1, a plumbing config) keeps ATen glue around our convs.
"""
import math

import os

import torch
from torch import nn
import torch.nn.functional as F

from . import _lib
from . import _tree
from . import conv as C
from . import reconstructor as _rec
from .reconstructor import conv2d

# ------------------------------------------------------------------------------------------------
PROGGAN_PLAN = ([(512, 512, 4, 3, False), (512, 512, 3, 1, False)]
                + [(512, 512, 3, 1, True), (512, 512, 3, 1, False)] * 3
                + [(512, 256, 3, 1, True), (256, 256, 3, 1, False), (256, 128, 3, 1, True), (128, 128, 3, 1, False),
                   (128, 64, 3, 1, True), (64, 64, 3, 1, False), (64, 32, 3, 1, True), (32, 32, 3, 1, False),
                   (32, 16, 3, 1, True), (16, 16, 3, 1, False)])


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


def _pixel_norm(x):
    return x / torch.sqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


class _Frozen(nn.Module):
    """Shared plumbing: CUDA-only, cached folded weights invalidated on load / device move."""

    def __init__(self):
        super().__init__()
        self._plan = None
        _rec._FROZEN_PACKS.clear()              # the pack cache is keyed by address: a new model may reuse freed memory

    def _apply(self, fn, *a, **k):
        self._plan = None
        _rec._FROZEN_PACKS.clear()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plan = None
        _rec._FROZEN_PACKS.clear()
        return super().load_state_dict(*a, **k)

    def _require_cuda(self, x):
        if not x.is_cuda:
            raise RuntimeError('%s runs on CUDA tensors only (no CPU fallback); got %s' % (type(self).__name__, x.device))


class ProgGANGenerator(_Frozen):
    """models/ProgGAN/model.py:65-95 as a hand-scheduled kernel chain (one autograd node, data-gradient only).

    Per block (reference: pixel-norm (3 kernels) -> nearest x2 -> cuDNN conv -> x*scale + b -> leaky_relu, :42-62):
      * `wgs_pixelnorm_pack`: previous activation -> pixel-norm -> split32 operand (one pass, csrc/proggan.cu);
      * tensor-core conv whose epilogue adds the bias and applies LeakyReLU(0.2) (act = 2); WScale's scale is folded into
        the frozen weights once;
      * nearest x2 + 3x3 conv = four output-phase 2x2 convs over the LOW-resolution operand with pre-summed taps: output
        row 2q+py reads up-sampled rows 2q+py-1 .. 2q+py+1 = source rows {q-1, q, q} (py = 0) or {q, q, q+1} (py = 1), so
        the three taps collapse to two per axis (w0 | w1+w2, resp. w0+w1 | w2) - 16 MACs per low-res pixel instead of 36.
    Backward: `wgs_pixelnorm_bwd_pack` (pixel-norm backward + LeakyReLU backward of the block below + operand pack) and
    one data-gradient conv per block; for up-sampling blocks the adjoint of the four phase convs is ONE 4x4 / stride-2
    conv over the high-resolution gradient."""

    def __init__(self, plan=None):
        super().__init__()
        self.blocks = list(plan or PROGGAN_PLAN)
        for i, (ci, co, k, _, _) in enumerate(self.blocks):
            bound = 1.0 / math.sqrt(ci * k * k)
            _tree.add(self, 'features.%d.conv.weight' % i, torch.empty(co, ci, k, k).uniform_(-bound, bound))
            _tree.add(self, 'features.%d.wscale.scale' % i, torch.randn(1))
            _tree.add(self, 'features.%d.wscale.b' % i, torch.randn(co))
        c = self.blocks[-1][1]
        _tree.add(self, 'output.conv.weight', torch.empty(3, c, 1, 1).uniform_(-1 / math.sqrt(c), 1 / math.sqrt(c)))
        _tree.add(self, 'output.wscale.scale', torch.randn(1))
        _tree.add(self, 'output.wscale.b', torch.randn(3))

    def plan(self):
        if self._plan is not None:
            return self._plan
        t = _tree.tensors(self)
        if t['output.conv.weight'].device.type != 'cuda':
            raise RuntimeError('ProgGANGenerator runs on CUDA only (no CPU fallback); call .cuda() first')
        P = []
        with torch.no_grad():
            for i, (ci, co, k, pad, up) in enumerate(self.blocks):
                n = 'features.%d' % i
                w = (t[n + '.conv.weight'] * t[n + '.wscale.scale']).detach().float()            # [co, ci, k, k]
                e = dict(ci=ci, co=co, k=k, pad=pad, up=up, bias=t[n + '.wscale.b'].detach().float().contiguous())
                if up:
                    assert k == 3 and pad == 1
                    e['w_fwd'], e['taps'], e['w_bwd'] = up_conv_weights(w)
                else:
                    e['w_fwd'] = C.pack_weights(w)
                    e['w_bwd'] = C.pack_weights(torch.flip(w, [2, 3]).permute(1, 0, 2, 3).contiguous())
                P.append(e)
            wo = (t['output.conv.weight'] * t['output.wscale.scale']).detach().float()             # [3, c, 1, 1]
            self._plan = dict(blocks=P, w_out=C.pack_weights(wo), w_out_bwd=C.pack_weights(wo.permute(1, 0, 2, 3).contiguous()),
                              b_out=t['output.wscale.b'].detach().float().contiguous(), c_out=wo.shape[1])
        return self._plan

    def synthesize_pair(self, x_plain, x_shifted):
        """Both images of a training pair in one batched pass: rows [x_plain; x_shifted], only the shifted rows are taped."""
        b = x_plain.shape[0]
        x_all = torch.cat([x_plain.detach().float(), x_shifted.float()], dim=0).contiguous()
        plain, shifted = _ProgGANPairFn.apply(self, x_all, b)          # NHWC halves of one buffer
        return plain.permute(0, 3, 1, 2), shifted.permute(0, 3, 1, 2)

    def forward(self, x):
        self._require_cuda(x)
        return _ProgGANFn.apply(self, x.reshape(x.shape[0], -1).float().contiguous(), 0).permute(0, 3, 1, 2)


def _pixelnorm_pack(a):
    """fp32 NHWC [N, H, W, C] -> split32 operand of pixel_norm(a) (models/ProgGAN/model.py:17-18)."""
    n, h, w, c = a.shape
    out = torch.empty(n, h, w, C.chunks_of(c), 64, device=a.device, dtype=torch.bfloat16)
    _lib.call('wgs_pixelnorm_pack', _lib.ptr(a), n * h * w, c, 1e-8, _lib.ptr(out), None, _lib.stream())
    return out


def _pixelnorm_bwd(dxn, a, slope, split):
    """Backward of pixel_norm at input `a` (+ LeakyReLU backward of the block that produced `a` when slope >= 0)."""
    n, h, w, c = a.shape
    if split:
        out = torch.empty(n, h, w, C.chunks_of(c), 64, device=a.device, dtype=torch.bfloat16)
        _lib.call('wgs_pixelnorm_bwd_pack', _lib.ptr(dxn), _lib.ptr(a), n * h * w, c, 1e-8, float(slope), _lib.ptr(out), None,
                  _lib.stream())
    else:
        out = torch.empty_like(a)
        _lib.call('wgs_pixelnorm_bwd_pack', _lib.ptr(dxn), _lib.ptr(a), n * h * w, c, 1e-8, float(slope), None, _lib.ptr(out),
                  _lib.stream())
    return out


# 1 = the conv epilogue emits pixel_norm(act) as the next block's split32 operand (wgs_conv_desc.pixnorm_eps) instead of a
# separate pixel-norm + pack pass.  Measured on one B200 (same call, B = 8): 23.5 / 23.3 ms fused vs 22.9 / 22.8 ms separate -
# the <= 64-channel convs at 256^2 .. 1024^2 are epilogue / shared-memory-pipe bound, so work moved INTO their epilogue costs
# more than a full-bandwidth pass of its own (the same outcome as BatchNorm statistics in the epilogue,
# profiles/r02_experiments.md).  Kept as a tested A/B switch.
FUSE_PIXNORM = os.environ.get('WGS_FUSE_PIXNORM', '0') == '1'


def _proggan_forward(G, x, tape, grad_from):
    """x [N, 512] -> image NHWC [N, H, W, 3]; tape (rows >= grad_from): the input activation of every block.

    With FUSE_PIXNORM, from 64 x 64 upwards (one image per 128-pixel tile, <= 256 channels) the conv epilogue itself emits
    the NEXT block's operand: pixel_norm(act) as split32 (wgs_conv_desc.pixnorm_eps) next to the fp32 activation of the
    back-propagated rows, instead of the separate pixel-norm + pack pass."""
    P = G.plan()
    n = x.shape[0]
    dev = x.device
    a = x.view(n, 1, 1, -1).contiguous()
    xs = _pixelnorm_pack(a)
    acts = []
    for e in P['blocks']:
        if tape is not None:
            acts.append(a[grad_from:])
        h, w = xs.shape[1], xs.shape[2]
        co = e['co']
        oh, ow = (2 * h, 2 * w) if e['up'] else (h + 2 * e['pad'] - e['k'] + 1, w + 2 * e['pad'] - e['k'] + 1)
        grid_px = h * w if e['up'] else oh * ow                        # pixels per image of one launch's output grid
        fuse = FUSE_PIXNORM and co <= 256 and co % 16 == 0 and grid_px >= 128
        need_f32 = (tape is not None) or not fuse
        a = torch.empty(n, oh, ow, co, device=dev, dtype=torch.float32) if need_f32 else None
        kw = dict(beta=e['bias'], act=2)
        if fuse:
            xs_next = torch.empty(n, oh, ow, C.chunks_of(co), 64, device=dev, dtype=torch.bfloat16)
            kw.update(out_split=xs_next, pixnorm_eps=1e-8, out_from_n=grad_from if tape is not None else 0)
        else:
            kw.update(split_k=2)
        if e['up']:
            up_conv_forward(xs, e['w_fwd'], e['taps'], co, e['ci'], out=a, out_n=n, **kw)
        else:
            C.conv2d(xs, e['w_fwd'], e['k'], e['k'], padding=e['pad'], out=a, no_f32=a is None, cin=e['ci'], **kw)
        xs = xs_next if fuse else _pixelnorm_pack(a)
    if tape is not None:
        acts.append(a[grad_from:])
        tape['acts'] = acts
    return C.conv2d(xs, P['w_out'], 1, 1, beta=P['b_out'], cin=P['c_out'])


def _proggan_backward(G, tape, dimg):
    """dimg NHWC [n, H, W, 3] -> d(loss)/dx [n, 512]."""
    P = G.plan()
    acts = tape['acts']
    dxn = C.conv2d(C.pack_split32(dimg.contiguous()), P['w_out_bwd'], 1, 1, cout=P['c_out'], cin=3)
    for i in range(len(P['blocks']) - 1, -1, -1):
        e = P['blocks'][i]
        gs = _pixelnorm_bwd(dxn, acts[i + 1], 0.2, split=True)      # gradient w.r.t. block i's pre-activation, packed
        if e['up']:
            dxn = C.conv2d(gs, e['w_bwd'], 4, 4, stride=2, padding=1, cout=e['ci'], cin=e['co'], split_k=2)
        else:
            dxn = C.conv2d(gs, e['w_bwd'], e['k'], e['k'], padding=e['k'] - 1 - e['pad'], cout=e['ci'], cin=e['co'], split_k=2)
    return _pixelnorm_bwd(dxn, acts[0], -1.0, split=False).reshape(dxn.shape[0], -1)


class _ProgGANFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, G, x_all, n_plain):
        need = ctx.needs_input_grad[1]
        tape = {} if need else None
        img = _proggan_forward(G, x_all.detach(), tape, n_plain)
        ctx.G, ctx.tape, ctx.n_plain, ctx.rows = G, tape, n_plain, x_all.shape[0]
        return img

    @staticmethod
    def backward(ctx, dimg):
        if ctx.tape is None:
            return None, None, None
        dx = _proggan_backward(ctx.G, ctx.tape, dimg[ctx.n_plain:])
        ctx.tape = None
        if ctx.n_plain == 0:
            return None, dx, None
        full = dx.new_zeros(ctx.rows, dx.shape[1])
        full[ctx.n_plain:] = dx
        return None, full, None


class _ProgGANPairFn(torch.autograd.Function):
    """Both halves of the batched pass as separate outputs: the gradient of the shifted half arrives as it is (no
    zero-padded full-batch gradient, no slice copies)."""

    @staticmethod
    def forward(ctx, G, x_all, n_plain):
        need = ctx.needs_input_grad[1]
        tape = {} if need else None
        img = _proggan_forward(G, x_all.detach(), tape, n_plain)
        ctx.G, ctx.tape, ctx.n_plain, ctx.rows = G, tape, n_plain, x_all.shape[0]
        plain, shifted = img[:n_plain], img[n_plain:]
        ctx.mark_non_differentiable(plain)
        return plain, shifted

    @staticmethod
    def backward(ctx, _dplain, dimg):
        if ctx.tape is None or dimg is None:
            return None, None, None
        dx = _proggan_backward(ctx.G, ctx.tape, dimg)
        ctx.tape = None
        full = dx.new_zeros(ctx.rows, dx.shape[1])
        full[ctx.n_plain:] = dx
        return None, full, None


# ------------------------------------------------------------------------------------------------
# nearest x2 up-sample folded into the following 3x3 conv: output phase p along one axis reads source rows
# {q-1, q, q} (p = 0) or {q, q, q+1} (p = 1), so the three taps collapse to two: (offset, kernel rows summed)
_UP_PHASE = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}
_UP_INV = ((1, 1), (0, 1), (1, 0), (0, 0))        # adjoint 4x4 / stride-2 kernel row ky <-> (phase, collapsed tap)


def up_conv_weights(w):
    """w [co, ci, 3, 3] (fp32) of a conv applied AFTER a nearest x2 up-sample -> (forward pack [16, co, ci] for four
    output-phase 2x2 convs over the low-resolution input, {phase: taps}, adjoint pack [16, ci, co] = ONE 4x4 / stride-2 /
    pad-1 conv over the high-resolution gradient)."""
    fw, taps = [], {}
    for py in range(2):
        for px in range(2):
            taps[(py, px)] = []
            for dy, kys in _UP_PHASE[py]:
                for dx, kxs in _UP_PHASE[px]:
                    taps[(py, px)].append((dy, dx, len(fw)))
                    fw.append(w[:, :, list(kys)][:, :, :, list(kxs)].sum(dim=(2, 3)))                  # [co, ci]
    bw = []
    for ky in range(4):
        for kx in range(4):
            (py, a), (px, b) = _UP_INV[ky], _UP_INV[kx]
            bw.append(fw[((py * 2 + px) * 2 + a) * 2 + b].t())                                          # [ci, co]
    return C.pack_weight_rows(torch.stack(fw).contiguous()), taps, C.pack_weight_rows(torch.stack(bw).contiguous())


def up_conv_forward(xs, w_fwd, taps, co, ci, out=None, **epilogue):
    """conv3x3(nearest_x2(x)) from the low-resolution split32 operand xs [N, h, w, chunks, 64] -> fp32 [N, 2h, 2w, co]
    (with a fused split32 output in `epilogue`, `out` may stay None when `out_n` is given)."""
    n, h, w = xs.shape[0], xs.shape[1], xs.shape[2]
    if out is None and epilogue.get('out_split') is None:
        out = torch.empty(n, 2 * h, 2 * w, co, device=xs.device, dtype=torch.float32)
    C.run_phases([lambda py=py, px=px, tp=tp: C.conv_taps(xs, w_fwd, tp, out, grid=(h, w), out_origin=(py, px), out_step=(2, 2),
                                                         cout=co, cin=ci, **epilogue)
                  for (py, px), tp in taps.items()], n * h * w * co, xs.device)
    return out


def affine_act_pack(x, A=None, B=None, relu=True, split=True):
    """[relu](A[n, c] * x + B[n, c]) of an fp32 NHWC tensor as the split32 operand of the next conv (csrc/ccbn.cu)."""
    n, h, w, c = x.shape
    shape = (n, h, w, C.chunks_of(c), 64) if split else x.shape
    out = torch.empty(shape, device=x.device, dtype=torch.bfloat16 if split else torch.float32)
    groups = A.shape[0] if A is not None else 1
    _lib.call('wgs_affine_act_pack', _lib.ptr(x), _lib.ptr(A), _lib.ptr(B), n * h * w, c, (n * h * w) // groups, int(relu),
              _lib.ptr(out) if split else None, None if split else _lib.ptr(out), _lib.stream())
    return out


def affine_act_bwd(dz, x, A, B, relu=True, split=True, want_sums=True, pad_rows=0):
    """-> (dx [split32 or fp32], dA, dB) for o = [relu](A x + B); A, B [groups, C] (groups = N, or 1 for a per-channel map).
    pad_rows > 0 (per-sample maps, fp32 dx): the results are the trailing rows of tensors with pad_rows leading ZERO rows -
    the gradient layout autograd wants for a batch whose first pad_rows samples carry no gradient, without a torch.cat."""
    n, h, w, c = x.shape
    groups = A.shape[0]
    assert pad_rows == 0 or (not split and (groups == n or not want_sums))
    if split:
        dx_full = torch.empty(n, h, w, C.chunks_of(c), 64, device=x.device, dtype=torch.bfloat16)
    elif pad_rows:
        dx_full = torch.zeros(pad_rows + n, h, w, c, device=x.device, dtype=torch.float32)
    else:
        dx_full = torch.empty(n, h, w, c, device=x.device, dtype=torch.float32)
    dx = dx_full[pad_rows:]
    sums_full = torch.zeros(2, pad_rows + groups, c, device=x.device, dtype=torch.float32) if want_sums else None
    sums = sums_full[:, pad_rows:] if want_sums else None
    _lib.call('wgs_affine_act_bwd', _lib.ptr(dz), _lib.ptr(x), _lib.ptr(A), _lib.ptr(B), groups, (n * h * w) // groups, c,
              int(relu), _lib.ptr(dx) if split else None, None if split else _lib.ptr(dx),
              _lib.ptr(sums[0]) if want_sums else None, _lib.ptr(sums[1]) if want_sums else None, _lib.stream())
    return (dx_full, sums_full[0], sums_full[1]) if want_sums else (dx_full, None, None)


# ------------------------------------------------------------------------------------------------
SN_RES_GEN_CONFIGS = {'sn_resnet32': ([256, 256, 256, 256], 4), 'sn_resnet64': ([1024, 512, 256, 128, 64], 4)}


class SNGANGenerator(_Frozen):
    """GenWrapper(model=Sequential(...)) of models/SNGAN/sn_gen_resnet.py:57-112; keys live under ``model.``."""

    def __init__(self, model='sn_resnet32', img_size=32, channels=1, latent_dim=128):
        super().__init__()
        self.cfg, self.seed = SN_RES_GEN_CONFIGS[model]
        self.model_name, self.img_size, self.image_channels, self.latent_dim = model, img_size, channels, latent_dim
        self.distribution = _tree.Node()
        self.distribution.dim = latent_dim
        ch, seed = self.cfg, self.seed

        def conv(p, ci, co):
            w = torch.empty(co, ci, 3, 3)
            nn.init.xavier_uniform_(w)
            _tree.add(self, p + '.weight', w)
            _tree.add(self, p + '.bias', torch.zeros(co))

        def bn(p, c):
            _tree.add(self, p + '.weight', torch.ones(c))
            _tree.add(self, p + '.bias', torch.zeros(c))
            _tree.add(self, p + '.running_mean', torch.zeros(c), buffer=True)
            _tree.add(self, p + '.running_var', torch.ones(c), buffer=True)
            _tree.add(self, p + '.num_batches_tracked', torch.tensor(0, dtype=torch.long), buffer=True)

        w0 = torch.empty(seed * seed * ch[0], latent_dim)
        nn.init.xavier_uniform_(w0)
        _tree.add(self, 'model.0.weight', w0)
        _tree.add(self, 'model.0.bias', torch.zeros(seed * seed * ch[0]))
        for i in range(len(ch) - 1):
            p = 'model.%d' % (2 + i)
            conv(p + '.conv1', ch[i], ch[i + 1])
            conv(p + '.conv2', ch[i + 1], ch[i + 1])
            bn(p + '.model.0', ch[i])
            bn(p + '.model.4', ch[i + 1])
            if ch[i] != ch[i + 1]:
                conv(p + '.bypass.1', ch[i], ch[i + 1])
        n = len(ch) + 1
        bn('model.%d' % n, ch[-1])
        conv('model.%d' % (n + 2), ch[-1], channels)

    def _bn(self, t, p, x, eps=1e-5):
        scale = t[p + '.weight'] * torch.rsqrt(t[p + '.running_var'] + eps)
        shift = t[p + '.bias'] - t[p + '.running_mean'] * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    def generate(self, z):
        self._require_cuda(z)
        t = {k: v.detach() for k, v in _tree.tensors(self).items()}
        ch, seed = self.cfg, self.seed
        x = F.linear(z, t['model.0.weight'], t['model.0.bias']).view(-1, ch[0], seed, seed)
        for i in range(len(ch) - 1):
            p = 'model.%d' % (2 + i)
            h = F.interpolate(F.relu(self._bn(t, p + '.model.0', x)), scale_factor=2, mode='nearest')
            h = conv2d(_cl(h), t[p + '.conv1.weight'], t[p + '.conv1.bias'], 1, 1)
            h = conv2d(_cl(F.relu(self._bn(t, p + '.model.4', h))), t[p + '.conv2.weight'], t[p + '.conv2.bias'], 1, 1)
            s = F.interpolate(x, scale_factor=2, mode='nearest')
            if ch[i] != ch[i + 1]:
                s = conv2d(_cl(s), t[p + '.bypass.1.weight'], t[p + '.bypass.1.bias'], 1, 1)
            x = h + s
        n = len(ch) + 1
        x = F.relu(self._bn(t, 'model.%d' % n, x))
        return torch.tanh(conv2d(_cl(x), t['model.%d.weight' % (n + 2)], t['model.%d.bias' % (n + 2)], 1, 1))

    def forward(self, z):
        return self.generate(z)


# ------------------------------------------------------------------------------------------------
def biggan_arch(resolution, ch=96, attention='64'):
    mult = {256: ([16, 16, 8, 8, 4, 2], [16, 8, 8, 4, 2, 1]), 128: ([16, 16, 8, 4, 2], [16, 8, 4, 2, 1]),
            64: ([16, 16, 8, 4], [16, 8, 4, 2]), 32: ([4, 4, 4], [4, 4, 4])}[resolution]
    attn = [int(a) for a in attention.split('_')]
    res = [8 * 2 ** i for i in range(len(mult[0]))]
    return {'in': [ch * m for m in mult[0]], 'out': [ch * m for m in mult[1]], 'attn': [r in attn for r in res]}


class _Embedding(_tree.Node):
    def forward(self, idx):
        return F.embedding(idx, self.weight.detach())


class BigGANGenerator(_Frozen):
    """models/BigGAN/BigGAN.py:54-243 with G_shared, hierarchical z, SN parametrisation, eval-mode BN."""

    def __init__(self, G_ch=96, dim_z=120, bottom_width=4, resolution=128, G_attn='64', n_classes=1000,
                 shared_dim=128, hier=True, BN_eps=1e-5, SN_eps=1e-6, **unused):
        super().__init__()
        self.ch, self.bottom_width, self.resolution, self.attention = G_ch, bottom_width, resolution, G_attn
        self.n_classes, self.shared_dim, self.hier, self.BN_eps, self.SN_eps = n_classes, shared_dim, hier, BN_eps, SN_eps
        self.arch = biggan_arch(resolution, G_ch, G_attn)
        nb = len(self.arch['out'])
        self.num_slots = nb + 1 if hier else 1
        self.z_chunk_size = dim_z // self.num_slots if hier else 0
        self.dim_z = self.z_chunk_size * self.num_slots if hier else dim_z
        cond = shared_dim + self.z_chunk_size

        def sn(p, shape, bias=True):
            _tree.add(self, p + '.weight', torch.randn(*shape) * 0.02)
            if bias:
                _tree.add(self, p + '.bias', torch.zeros(shape[0]))
            _tree.add(self, p + '.u0', torch.randn(1, shape[0]), buffer=True)
            _tree.add(self, p + '.sv0', torch.ones(1), buffer=True)

        def stats(p, c):
            _tree.add(self, p + '.stored_mean', torch.zeros(c), buffer=True)
            _tree.add(self, p + '.stored_var', torch.ones(c), buffer=True)

        self.shared = _Embedding()
        _tree.add(self, 'shared.weight', torch.randn(n_classes, shared_dim))
        sn('linear', (self.arch['in'][0] * bottom_width ** 2, self.dim_z // self.num_slots))
        for i in range(nb):
            p = 'blocks.%d.0' % i
            ci, co = self.arch['in'][i], self.arch['out'][i]
            sn(p + '.conv1', (co, ci, 3, 3))
            sn(p + '.conv2', (co, co, 3, 3))
            sn(p + '.conv_sc', (co, ci, 1, 1))
            for name, c in (('.bn1', ci), ('.bn2', co)):
                sn(p + name + '.gain', (c, cond), bias=False)
                sn(p + name + '.bias', (c, cond), bias=False)
                stats(p + name, c)
            if self.arch['attn'][i]:
                q = 'blocks.%d.1' % i
                sn(q + '.theta', (co // 8, co, 1, 1), bias=False)
                sn(q + '.phi', (co // 8, co, 1, 1), bias=False)
                sn(q + '.g', (co // 2, co, 1, 1), bias=False)
                sn(q + '.o', (co, co // 2, 1, 1), bias=False)
                _tree.add(self, q + '.gamma', torch.tensor(0.))
        c = self.arch['out'][-1]
        _tree.add(self, 'output_layer.0.gain', torch.ones(c))
        _tree.add(self, 'output_layer.0.bias', torch.zeros(c))
        stats('output_layer.0', c)
        sn('output_layer.2', (3, c, 3, 3))

    def plan(self):
        """Frozen, kernel-ready state: W / sigma for every spectrally-normalised layer (one power iteration off the stored
        u0, exactly what the reference's eval forward redoes on every call, models/BigGAN/layers.py:84-96), packed conv
        weights (up-sampling convs as collapsed phase taps) and the eval-BatchNorm constants."""
        if self._plan is not None:
            return self._plan
        t = {k: v.detach() for k, v in _tree.tensors(self).items()}
        if t['shared.weight'].device.type != 'cuda':
            raise RuntimeError('BigGANGenerator runs on CUDA only (no CPU fallback); call .cuda() first')
        sn = {}
        with torch.no_grad():
            for k in t:
                if k.endswith('.u0'):
                    p = k[:-3]
                    w = t[p + '.weight']
                    wm = w.reshape(w.shape[0], -1)
                    v = F.normalize(t[k] @ wm, eps=self.SN_eps)
                    u2 = F.normalize(v @ wm.t(), eps=self.SN_eps)
                    sn[p] = (w / torch.squeeze((v @ wm.t()) @ u2.t())).float().contiguous()
            blocks = []
            for i in range(len(self.arch['out'])):
                p = 'blocks.%d.0' % i
                ci, co = self.arch['in'][i], self.arch['out'][i]
                e = dict(ci=ci, co=co)
                e['w1'], e['taps1'], e['w1_bwd'] = up_conv_weights(sn[p + '.conv1'])
                e['w2'] = C.pack_weights(sn[p + '.conv2'])
                e['w2_bwd'] = C.pack_weights(torch.flip(sn[p + '.conv2'], [2, 3]).permute(1, 0, 2, 3).contiguous())
                e['wsc'] = C.pack_weights(sn[p + '.conv_sc'])
                e['wsc_bwd'] = C.pack_weights(sn[p + '.conv_sc'].permute(1, 0, 2, 3).contiguous())
                for name in ('conv1', 'conv2', 'conv_sc'):
                    e['b_' + name] = t['%s.%s.bias' % (p, name)].float().contiguous()
                for bn in ('bn1', 'bn2'):
                    # ccbn in eval mode is affine in the class / latent vector y (layers.py:303-322): with inv = rsqrt(var + eps)
                    #   A = (1 + y Wg^T) inv = inv + y (Wg inv)^T,   B = y Wb^T - (1 + y Wg^T) mean inv = -mean inv + y (Wb - Wg mean inv)^T
                    # so the constants are folded into the two linears once and each of A / B is ONE addmm per call (it was two
                    # matrix-vector products and four element-wise kernels per BatchNorm, ~270 launches per step at config 4)
                    inv = torch.rsqrt(t['%s.%s.stored_var' % (p, bn)] + self.BN_eps).float()
                    mean_inv = (t['%s.%s.stored_mean' % (p, bn)] * inv).float()
                    wg, wb = sn['%s.%s.gain' % (p, bn)], sn['%s.%s.bias' % (p, bn)]
                    e[bn] = ((wg * inv[:, None]).t().contiguous(), inv.contiguous(),
                             (wb - wg * mean_inv[:, None]).t().contiguous(), (-mean_inv).contiguous())
                blocks.append(e)
            p = 'output_layer.0'
            a = t[p + '.gain'] * torch.rsqrt(t[p + '.stored_var'] + self.BN_eps)
            wo = sn['output_layer.2']
            out = dict(A=a.float().reshape(1, -1).contiguous(), B=(t[p + '.bias'] - t[p + '.stored_mean'] * a).float().reshape(1, -1).contiguous(),
                       w=C.pack_weights(wo), w_bwd=C.pack_weights(torch.flip(wo, [2, 3]).permute(1, 0, 2, 3).contiguous()),
                       bias=t['output_layer.2.bias'].float().contiguous(), ci=wo.shape[1])
        self._plan = dict(sn=sn, blocks=blocks, out=out)
        return self._plan

    def _ccbn_affine(self, sn, e, p, bn, y):
        """ccbn in eval mode as a per-sample affine map (layers.py:303-322): A = gain / sqrt(var + eps), B = bias - mean * A."""
        wa_t, ba, wb_t, bb = e[bn]
        return torch.addmm(ba, y, wa_t), torch.addmm(bb, y, wb_t)

    def forward(self, z, y, grad_from=0):
        """-> logical NCHW image (channels-last memory).  models/BigGAN/BigGAN.py:222-243.
        grad_from > 0 (the batched pair pass): rows [0, grad_from) are the un-shifted images, which need no gradient - the
        hand-scheduled nodes back-propagate rows >= grad_from only."""
        self._require_cuda(z)
        P = self.plan()
        sn = P['sn']
        t = _tree.tensors(self)
        nb = len(self.arch['out'])
        if self.hier:
            zs = torch.split(z, self.z_chunk_size, 1)
            z = zs[0]
            ys = [torch.cat([y, item], 1) for item in zs[1:]]
        else:
            ys = [y] * nb
        h = F.linear(z, sn['linear'], t['linear.bias'].detach()).view(z.shape[0], -1, self.bottom_width, self.bottom_width)
        h = h.permute(0, 2, 3, 1).contiguous()                                  # NHWC from here on
        # The ccbn maps of every block depend on (y, z) alone: all of them are launched up front on a forked stream (four small
        # library GEMMs per block) and each block waits for its own four only - off the critical path from block 1 on, forward
        # and (autograd runs a node's backward on its forward stream) backward.
        main = torch.cuda.current_stream()
        aux = C._phase_streams(z.device, 5)[4] if (CCBN_FORK and C.PROFILE is None) else None
        if aux is not None:
            fork = torch.cuda.Event()
            fork.record(main)
            aux.wait_event(fork)
        ab = []
        with torch.cuda.stream(aux if aux is not None else main):
            for i, e in enumerate(P['blocks']):
                p = 'blocks.%d.0' % i
                maps = self._ccbn_affine(sn, e, p, 'bn1', ys[i]) + self._ccbn_affine(sn, e, p, 'bn2', ys[i])
                ready = None
                if aux is not None:
                    ready = torch.cuda.Event()
                    ready.record(aux)
                ab.append((maps, ready))
        for i, e in enumerate(P['blocks']):
            (A1, B1, A2, B2), ready = ab[i]
            if ready is not None:
                main.wait_event(ready)
            h = _GBlockFn.apply(h, A1, B1, A2, B2, e, grad_from)
            if self.arch['attn'][i]:
                if grad_from > 0:          # library ops under autograd: keep the un-shifted rows out of the graph
                    # (and on the forked stream: the two halves are independent chains of small library kernels)
                    if aux is not None:
                        fork = torch.cuda.Event()
                        fork.record(main)
                        aux.wait_event(fork)
                    with torch.cuda.stream(aux if aux is not None else main), torch.no_grad():
                        h_plain = self._attention(sn, t, 'blocks.%d.1' % i, h[:grad_from])
                        if aux is not None:
                            plain_done = torch.cuda.Event()
                            plain_done.record(aux)
                    h_shifted = self._attention(sn, t, 'blocks.%d.1' % i, h[grad_from:])
                    if aux is not None:
                        main.wait_event(plain_done)
                    h = torch.cat([h_plain, h_shifted], dim=0)
                else:
                    h = self._attention(sn, t, 'blocks.%d.1' % i, h)
        return _OutputFn.apply(h, P['out'], grad_from).permute(0, 3, 1, 2)

    def synthesize_pair(self, z_plain, y_plain, z_shifted, y_shifted):
        """Both images of a training pair in ONE batched pass (rows [plain; shifted]); only the shifted rows are
        back-propagated.  The 4 x 4 .. 32 x 32 blocks are latency-bound: two passes of B images cost twice one pass of 2B."""
        b = z_plain.shape[0]
        img = self.forward(torch.cat([z_plain.detach(), z_shifted], dim=0), torch.cat([y_plain.detach(), y_shifted], dim=0),
                           grad_from=b)
        return img[:b].detach(), img[b:]

    def _attention(self, sn, t, q, h):
        """layers.py:153-166.  The four 1x1 convs run on the tensor-core kernel; the two batched matrix products
        (4096 x 24 x 1024 and 96 x 1024 x 4096 per sample, ~1 GFLOP = 2 % of the generator), soft-max and 2x2 max-pool are
        library calls (cuBLAS bmm / ATen) under autograd."""
        x = h.permute(0, 3, 1, 2)                                               # logical NCHW view of NHWC memory
        b, c, hh, ww = x.shape
        theta = conv2d(x, sn[q + '.theta'], None, 1, 0).reshape(b, c // 8, hh * ww)
        phi = F.max_pool2d(conv2d(x, sn[q + '.phi'], None, 1, 0), [2, 2]).reshape(b, c // 8, hh * ww // 4)
        g = F.max_pool2d(conv2d(x, sn[q + '.g'], None, 1, 0), [2, 2]).reshape(b, c // 2, hh * ww // 4)
        beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
        o = torch.bmm(g, beta.transpose(1, 2)).view(b, c // 2, hh, ww)
        out = t[q + '.gamma'].detach() * conv2d(_cl(o), sn[q + '.o'], None, 1, 0) + x
        return out.permute(0, 2, 3, 1).contiguous()


def _affine_bwd_split_padded(dz, x, A, B, pad_rows):
    """affine_act_bwd with a split32 dx and dA / dB carrying pad_rows leading zero rows."""
    n, h, w, c = x.shape
    dx = torch.empty(n, h, w, C.chunks_of(c), 64, device=x.device, dtype=torch.bfloat16)
    sums = torch.zeros(2, pad_rows + n, c, device=x.device, dtype=torch.float32)
    _lib.call('wgs_affine_act_bwd', _lib.ptr(dz), _lib.ptr(x), _lib.ptr(A), _lib.ptr(B), n, h * w, c, 1, _lib.ptr(dx), None,
              _lib.ptr(sums[0, pad_rows:]), _lib.ptr(sums[1, pad_rows:]), _lib.stream())
    return dx, sums[0], sums[1]


CCBN_FORK = os.environ.get('WGS_CCBN_FORK', '1') != '0'        # 0 = BigGAN's ccbn linears in line, in front of each block (A/B switch)
GBLOCK_FORK = os.environ.get('WGS_GBLOCK_FORK', '1') != '0'      # 0 = BigGAN's shortcut branch in line with the main branch (A/B switch)


class _GBlockFn(torch.autograd.Function):
    """GBlock (models/BigGAN/layers.py:395-405) as one hand-scheduled node over NHWC fp32 activations:
        r1 = relu(A1 x + B1) -> conv1(up(r1)) + b1 = h1 -> r2 = relu(A2 h1 + B2) -> conv2(r2) + b2 + up(conv_sc(x) + b_sc)
    six tensor-core launch groups + three element-wise passes instead of ~14 full-tensor ATen passes around cuDNN convs:
      * ccbn + ReLU + operand pack is ONE pass (wgs_affine_act_pack), the up-sample never materialises (phase convs);
      * the 1x1 shortcut is computed at LOW resolution (1x1 conv commutes with nearest up-sampling) and written straight
        into the four output phases; conv2 accumulates on top of it in its epilogue.
    Backward carries the data gradient and the per-sample dA / dB (which autograd takes on to z through the ccbn linears)."""

    @staticmethod
    def forward(ctx, x, A1, B1, A2, B2, e, grad_from=0):
        x = x.contiguous()
        n, h, w, ci = x.shape
        co = e['co']
        # (split_k = 2: batch-independent cluster split-K for the 4 x 4 .. 16 x 16 blocks - 1536 .. 384 channels, up to 432
        # dependent contraction steps per CTA otherwise)
        # The shortcut branch (operand pack + 1x1 up-conv) depends on x alone: it runs on a forked stream beside
        # ccbn -> conv1 -> ccbn of the main branch and is joined in front of conv2, which accumulates on top of it.
        main = torch.cuda.current_stream()
        aux = C._phase_streams(x.device, 4)[3] if (GBLOCK_FORK and C.PROFILE is None) else None
        if aux is not None:
            fork = torch.cuda.Event()
            fork.record(main)
            aux.wait_event(fork)
        with torch.cuda.stream(aux if aux is not None else main):
            xs = affine_act_pack(x, relu=False)
            out = up_conv_forward(xs, e['wsc'], {(py, px): [(0, 0, 0)] for py in range(2) for px in range(2)}, co, ci,
                                  beta=e['b_conv_sc'], split_k=2)
            if aux is not None:
                joined = torch.cuda.Event()
                joined.record(aux)
        r1s = affine_act_pack(x, A1.detach(), B1.detach(), relu=True)
        h1 = up_conv_forward(r1s, e['w1'], e['taps1'], co, ci, beta=e['b_conv1'], split_k=2)
        r2s = affine_act_pack(h1, A2.detach(), B2.detach(), relu=True)
        if aux is not None:
            main.wait_event(joined)
        C.conv2d(r2s, e['w2'], 3, 3, padding=1, out=out, accumulate=True, beta=e['b_conv2'], cin=co, split_k=2)
        if any(ctx.needs_input_grad[:5]):
            g0 = grad_from
            ctx.save_for_backward(x[g0:], h1[g0:], A1[g0:], B1[g0:], A2[g0:], B2[g0:])
        ctx.e, ctx.g0, ctx.rows = e, grad_from, n
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, h1, A1, B1, A2, B2 = ctx.saved_tensors           # rows >= grad_from only
        e, g0 = ctx.e, ctx.g0
        # (rows < grad_from carry no gradient: the kernels write rows >= grad_from of zero-initialised full-batch tensors)
        return _GBlockFn._backward_rows(x, h1, A1, B1, A2, B2, e, d_out[g0:] if g0 else d_out, g0) + (None, None)

    @staticmethod
    def _backward_rows(x, h1, A1, B1, A2, B2, e, d_out, pad_rows=0):
        n, h, w, ci = x.shape
        co = e['co']
        ds = C.pack_split32(d_out.contiguous())
        # shortcut: adjoint of up(conv_sc(x)) = conv_sc^T of the 2x2 sum-pooled gradient = four stride-2 taps of the 1x1 weight;
        # it needs ds alone, so it runs on the forked stream beside the main chain and is added at the end
        sc_taps = [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0)]
        main = torch.cuda.current_stream()
        aux = C._phase_streams(x.device, 4)[3] if (GBLOCK_FORK and C.PROFILE is None) else None
        if aux is not None:
            dsc = torch.empty(n, h, w, ci, device=x.device, dtype=torch.float32)
            fork = torch.cuda.Event()
            fork.record(main)
            aux.wait_event(fork)
            with torch.cuda.stream(aux):
                C.conv_taps(ds, e['wsc_bwd'], sc_taps, dsc, grid=(h, w), in_stride=2, cout=ci, cin=co, split_k=2)
                joined = torch.cuda.Event()
                joined.record(aux)
        dr2 = C.conv2d(ds, e['w2_bwd'], 3, 3, padding=1, cout=co, cin=co, split_k=2)
        dh1s, dA2, dB2 = _affine_bwd_split_padded(dr2, h1, A2.detach(), B2.detach(), pad_rows)
        dr1 = C.conv2d(dh1s, e['w1_bwd'], 4, 4, stride=2, padding=1, cout=ci, cin=co, split_k=2)     # adjoint of up + conv1
        dx_full, dA1, dB1 = affine_act_bwd(dr1, x, A1.detach(), B1.detach(), relu=True, split=False, pad_rows=pad_rows)
        dx = dx_full[pad_rows:]
        if aux is not None:
            main.wait_event(joined)
            dx.add_(dsc)
        else:
            C.conv_taps(ds, e['wsc_bwd'], sc_taps, dx, grid=(h, w), in_stride=2, cout=ci, cin=co, accumulate=True, split_k=2)
        return dx_full, dA1, dB1, dA2, dB2


class _OutputFn(torch.autograd.Function):
    """output_layer: eval BatchNorm -> ReLU -> SNConv 3x3 -> tanh (models/BigGAN/BigGAN.py:200-203,242-243): one affine
    pack pass + one conv with bias and tanh in its epilogue."""

    @staticmethod
    def forward(ctx, h, o, grad_from=0):
        h = h.contiguous()
        img = C.conv2d(affine_act_pack(h, o['A'], o['B'], relu=True), o['w'], 3, 3, padding=1, beta=o['bias'], act=4, cin=o['ci'])
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(h[grad_from:], img[grad_from:])
        ctx.o, ctx.g0 = o, grad_from
        return img

    @staticmethod
    def backward(ctx, dimg):
        h, img = ctx.saved_tensors                                              # rows >= grad_from only
        o, g0 = ctx.o, ctx.g0
        g = (dimg[g0:] * (1.0 - img * img)).contiguous()                        # tanh'
        dr = C.conv2d(C.pack_split32(g), o['w_bwd'], 3, 3, padding=1, cout=o['ci'], cin=3)
        dh, _, _ = affine_act_bwd(dr, h, o['A'], o['B'], relu=True, split=False, want_sums=False, pad_rows=g0)
        return dh, None, None
