"""StyleGAN2 generator on the libwgs_b200 kernels.

Mirrors the reference ``models.StyleGAN2.model.Generator`` (models/StyleGAN2/model.py:285-408): same
constructor arguments, attributes (``size``, ``style_dim``, ``n_latent``, ``num_layers``, ``channels``),
``get_latent`` / ``forward(styles, input_is_latent=...) -> (image, None)`` and the same 171 state-dict
keys, so a reference ``g_ema`` checkpoint loads with ``load_state_dict``.

B200-first restatement of the math (verified to 1e-15 in fp64):
  * modulated conv  y[b,o] = d[b,o] * conv(W, s[b,i] * x[b,i])  with  d = rsqrt(scale^2 * sum_i s^2 * Wsq[o,i]
    + 1e-8), Wsq = sum_k W^2 — the per-sample [B,Co,Ci,k,k] weight tensor and the groups=B conv of
    model.py:190-226 disappear; the whole batch is one dense implicit GEMM on tensor cores;
  * the stride-2 transposed conv (:201-206) is four output-phase tap lists of the same kernel;
  * blur + demod + noise + bias + sqrt(2)*lrelu is one FIR kernel; same-resolution layers fuse
    demod + noise + bias + activation into the conv epilogue; ToRGB + skip upsample is one kernel.
The frozen generator receives no weight gradients; backward only carries the data gradient to the styles.
"""
import ctypes
import os
import math

import torch
from torch import nn

from . import _lib
from . import conv as C


def default_channels(channel_multiplier=2):
    c = {4: 512, 8: 512, 16: 512, 32: 512}
    for res, base in ((64, 256), (128, 128), (256, 64), (512, 32), (1024, 16)):
        c[res] = base * channel_multiplier
    return c


def _fir2d(taps=(1, 3, 3, 1), gain=1.0):
    k = torch.tensor(taps, dtype=torch.float32)
    k = torch.outer(k, k)
    return k / k.sum() * gain


class _Node(nn.Module):
    """Bare container used to reproduce the reference module tree (and therefore its state-dict keys)."""


def _equal_linear(in_dim, out_dim, bias_init=0.0, lr_mul=1.0):
    m = _Node()
    m.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
    m.bias = nn.Parameter(torch.full((out_dim,), float(bias_init)))
    return m


def _mod_conv(ci, co, k, style_dim, upsample=False):
    m = _Node()
    m.weight = nn.Parameter(torch.randn(1, co, ci, k, k))
    m.modulation = _equal_linear(style_dim, ci, bias_init=1.0)
    if upsample:
        m.blur = _Node()
        m.blur.register_buffer('kernel', _fir2d(gain=4.0))
    return m


def _styled_conv(ci, co, style_dim, upsample=False):
    m = _Node()
    m.conv = _mod_conv(ci, co, 3, style_dim, upsample)
    m.noise = _Node()
    m.noise.weight = nn.Parameter(torch.zeros(1))
    m.activate = _Node()
    m.activate.bias = nn.Parameter(torch.zeros(co))
    return m


def _to_rgb(ci, style_dim, upsample=True):
    m = _Node()
    if upsample:
        m.upsample = _Node()
        m.upsample.register_buffer('kernel', _fir2d(gain=4.0))
    m.conv = _mod_conv(ci, 3, 1, style_dim)
    m.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))
    return m


MERGE_PHASES = os.environ.get('WGS_MERGE_PHASES', '1') != '0'
MERGE_WIDE = os.environ.get('WGS_MERGE_WIDE', '0') != '0'      # 1 = also the wide (> 64 channel) up-convs as one phase-packed launch (measured 1 % slower)
_TAPS = (ctypes.c_float * 4)(0.25, 0.75, 0.75, 0.25)          # [1,3,3,1]/8 * 2 per axis (kernel * 4 overall)


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), lr_mlp=0.01,
                 channels=None):
        super().__init__()
        if tuple(blur_kernel) != (1, 3, 3, 1):
            raise NotImplementedError('only the [1,3,3,1] FIR of the released models is implemented')
        self.size, self.style_dim, self.n_mlp, self.lr_mlp = size, style_dim, n_mlp, lr_mlp
        self.channels = dict(channels) if channels else default_channels(channel_multiplier)
        self.log_size = int(math.log2(size))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.n_latent = self.log_size * 2 - 2
        self.style = nn.ModuleList([_Node()] + [_equal_linear(style_dim, style_dim, lr_mul=lr_mlp)
                                                for _ in range(n_mlp)])
        self.input = _Node()
        self.input.input = nn.Parameter(torch.randn(1, self.channels[4], 4, 4))
        self.conv1 = _styled_conv(self.channels[4], self.channels[4], style_dim)
        self.to_rgb1 = _to_rgb(self.channels[4], style_dim, upsample=False)
        self.convs, self.to_rgbs = nn.ModuleList(), nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.noises = _Node()
        for layer in range(self.num_layers):
            res = (layer + 5) // 2
            self.noises.register_buffer('noise_%d' % layer, torch.randn(1, 1, 2 ** res, 2 ** res))
        cin = self.channels[4]
        for i in range(3, self.log_size + 1):
            co = self.channels[2 ** i]
            self.convs.append(_styled_conv(cin, co, style_dim, upsample=True))
            self.convs.append(_styled_conv(co, co, style_dim))
            self.to_rgbs.append(_to_rgb(co, style_dim))
            cin = co
        self._plan = None

    # ------------------------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._plan = None                       # packed weights live on the old device
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def plan(self):
        """Kernel-ready (frozen) weights: split32 packs, squared-weight sums, concatenated modulation."""
        if self._plan is not None:
            return self._plan
        dev = self.input.input.device
        if dev.type != 'cuda':
            raise RuntimeError('StyleGAN2 Generator runs on CUDA only (no CPU fallback); call .cuda() first')
        P = {}
        layers = [('conv1', self.conv1, False)]
        for i in range(self.log_size - 2):
            layers.append(('convs.%d' % (2 * i), self.convs[2 * i], True))
            layers.append(('convs.%d' % (2 * i + 1), self.convs[2 * i + 1], False))
        rgbs = [self.to_rgb1] + list(self.to_rgbs)
        mod_w, mod_b, offs = [], [], []
        off = 0
        with torch.no_grad():
            P['styled'] = []
            for name, m, up in layers:
                w = m.conv.weight[0].float()                                  # [Co, Ci, 3, 3]
                co, ci = w.shape[:2]
                scale = 1.0 / math.sqrt(ci * 9)
                ws = w * scale                                                # equalised-lr scale folded in
                ent = dict(name=name, up=up, ci=ci, co=co, scale=scale, s_off=off,
                           w_fwd=C.pack_weights(ws),
                           wsq=w.pow(2).sum(dim=[2, 3]).contiguous(),          # [Co, Ci]
                           noise_w=float(m.noise.weight.item()), bias=m.activate.bias.detach().float().contiguous())
                # data-gradient weights: dx[ci] = sum_{taps,co} dy[co] * W[co,ci,tap]
                if up:
                    ent['w_bwd'] = C.pack_weights(ws.permute(1, 0, 2, 3).contiguous())         # strided conv, same taps
                    if MERGE_PHASES and (co <= 64 or MERGE_WIDE):
                        # all four output phases of the transposed conv stacked along N (one launch, conv.py).  For the wide
                        # layers the 16/9 extra MACs of the zero blocks cost more than three launches save (A/B on one
                        # B200: 16.61 / 16.71 ms packed vs 16.52 / 16.44 ms as four phase launches): opt-in
                        shifts, idx, G = C._phase_plan('convT', 3, 3, 2, 0, dev)
                        ent['w_up'] = C.merged_phase_weights(ws.reshape(co, ci, 9).contiguous(), idx, len(shifts), G)
                else:
                    ent['w_bwd'] = C.pack_weights(torch.flip(ws, [2, 3]).permute(1, 0, 2, 3).contiguous())
                ent['wsq_t'] = ent['wsq'].t().contiguous()                    # [Ci, Co]
                mod_w.append(m.conv.modulation.weight.detach().float())
                mod_b.append(m.conv.modulation.bias.detach().float())
                off += ci
                P['styled'].append(ent)
            P['rgb'] = []
            for m in rgbs:
                w = m.conv.weight[0, :, :, 0, 0].float().contiguous()          # [3, C]
                ci = w.shape[1]
                P['rgb'].append(dict(ci=ci, scale=1.0 / math.sqrt(ci), s_off=off, w=w,
                                     bias=m.bias.detach().float().reshape(3).contiguous()))
                mod_w.append(m.conv.modulation.weight.detach().float())
                mod_b.append(m.conv.modulation.bias.detach().float())
                off += ci
            P['mod_w'] = torch.cat(mod_w, 0).contiguous()                     # [sumC, style_dim]
            P['mod_b'] = torch.cat(mod_b, 0).contiguous()
            P['mod_w_t'] = P['mod_w'].t().contiguous()                        # [style_dim, sumC]
            P['sum_c'] = off
            P['map_w'] = [self.style[i].weight.detach().float().contiguous() for i in range(1, self.n_mlp + 1)]
            P['map_b'] = [self.style[i].bias.detach().float().contiguous() for i in range(1, self.n_mlp + 1)]
            P['map_w_t'] = [w.t().contiguous() for w in P['map_w']]
            P['const'] = self.input.input[0].detach().float().permute(1, 2, 0).contiguous()   # [4,4,C] NHWC
            P['noise'] = [getattr(self.noises, 'noise_%d' % i)[0, 0].detach().float().contiguous()
                          for i in range(self.num_layers)]
            d_total = sum(e['co'] for e in P['styled'])
            P['d_total'] = d_total
        self._plan = P
        return P

    # ------------------------------------------------------------------------------------------
    def get_latent(self, z):
        """Mapping network: PixelNorm + n_mlp x (EqualLinear, lr_mul, fused lrelu) (model.py:291-295)."""
        return _MappingFn.apply(self, z.float().contiguous())

    def synthesize_pair(self, w_plain, w_shifted):
        """Both images of a training pair in ONE batched pass: rows [w_plain; w_shifted]; only the shifted
        rows are taped / back-propagated.  Returns (img_plain, img_shifted) as logical NCHW."""
        b = w_plain.shape[0]
        w_all = torch.cat([w_plain.detach().float(), w_shifted.float()], dim=0).contiguous()
        img_plain, img_shifted = _PairFn.apply(self, w_all, b)                 # NHWC halves of one buffer
        return img_plain.permute(0, 3, 1, 2), img_shifted.permute(0, 3, 1, 2)

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=False):
        if len(styles) != 1 or styles[0].dim() != 2 or truncation != 1 or noise is not None or randomize_noise:
            raise NotImplementedError('hot path only: one [B, style_dim] latent, fixed noise buffers, no truncation')
        x = styles[0].float().contiguous()
        w = x if input_is_latent else self.get_latent(x)
        img = _SynthesisFn.apply(self, w)                                     # NHWC
        image = img.permute(0, 3, 1, 2)                                       # logical NCHW, channels-last memory
        if return_latents:
            return image, w.unsqueeze(1).repeat(1, self.n_latent, 1)
        return image, None


# ----------------------------------------------------------------------------------------------
def _linear(x, W, bias, out, *, wscale=1.0, bscale=1.0, in_square=0, epi=0, eps=0.0, accumulate=0):
    B, I = x.shape
    O = W.shape[0]
    assert x.stride(1) == 1 and W.stride(1) == 1 and out.stride(1) == 1
    _lib.call('wgs_linear_small', ctypes.c_void_p(x.data_ptr()), x.stride(0), ctypes.c_void_p(W.data_ptr()), W.stride(0),
              _lib.ptr(bias), ctypes.c_void_p(out.data_ptr()), out.stride(0), B, I, O, float(wscale), float(bscale),
              int(in_square), int(epi), float(eps), int(accumulate), _lib.stream())
    return out


class LinearProblem(ctypes.Structure):
    """Mirror of wgs_linear_problem (include/wgs_b200.h)."""
    _fields_ = [('x', ctypes.c_void_p), ('x_ld', ctypes.c_longlong), ('x2', ctypes.c_void_p), ('x2_ld', ctypes.c_longlong),
                ('W', ctypes.c_void_p), ('w_ld', ctypes.c_longlong), ('bias', ctypes.c_void_p),
                ('mul', ctypes.c_void_p), ('mul_ld', ctypes.c_longlong), ('out', ctypes.c_void_p), ('out_ld', ctypes.c_longlong),
                ('I', ctypes.c_int), ('O', ctypes.c_int), ('wscale', ctypes.c_float), ('bscale', ctypes.c_float),
                ('eps', ctypes.c_float), ('in_mode', ctypes.c_int), ('epi', ctypes.c_int), ('accumulate', ctypes.c_int)]


_MAX_GROUP = 32
_layout_checked = False


def _problem(x, W, out, *, x2=None, bias=None, mul=None, wscale=1.0, bscale=1.0, eps=0.0, in_mode=0, epi=0, accumulate=0):
    """One small linear out[b,o] (+)= mul * epi(wscale * sum_i f(x, x2)[b,i] W[o,i] + bscale * bias[o]) (row-strided views ok)."""
    for t in (x, W, out, x2, mul):
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError('StyleGAN2 kernels need CUDA tensors (no CPU fallback); got %s' % t.device)
            assert t.stride(-1) == 1
    q = LinearProblem()
    q.x, q.x_ld = x.data_ptr(), x.stride(0)
    q.x2, q.x2_ld = (x2.data_ptr(), x2.stride(0)) if x2 is not None else (None, 0)
    q.W, q.w_ld = W.data_ptr(), W.stride(0)
    q.bias = bias.data_ptr() if bias is not None else None
    q.mul, q.mul_ld = (mul.data_ptr(), mul.stride(0)) if mul is not None else (None, 0)
    q.out, q.out_ld = out.data_ptr(), out.stride(0)
    q.I, q.O = x.shape[1], W.shape[0]
    q.wscale, q.bscale, q.eps = float(wscale), float(bscale), float(eps)
    q.in_mode, q.epi, q.accumulate = int(in_mode), int(epi), int(accumulate)
    return q


def _linear_group(problems, B):
    """All `problems` (same batch size B) in as few launches as the C ABI's group limit allows (one, here)."""
    global _layout_checked
    lib = _lib.load()
    if not _layout_checked:
        if lib.wgs_linear_problem_size() != ctypes.sizeof(LinearProblem):
            raise RuntimeError('wgs_linear_problem layout mismatch between header and ctypes mirror')
        _layout_checked = True
    for lo in range(0, len(problems), _MAX_GROUP):
        chunk = problems[lo: lo + _MAX_GROUP]
        arr = (LinearProblem * len(chunk))(*chunk)
        _lib.check(lib.wgs_linear_group(arr, len(chunk), int(B), _lib.stream()))


class MlpLayer(ctypes.Structure):
    """Mirror of wgs_mlp_layer (include/wgs_b200.h)."""
    _fields_ = [('W', ctypes.c_void_p), ('bias', ctypes.c_void_p), ('aux', ctypes.c_void_p), ('out', ctypes.c_void_p)]


MLP_CHAIN = os.environ.get('WGS_MLP_CHAIN', '1') != '0'      # 0 = one launch per mapping layer (A/B switch)
_mlp_checked = False


def _mlp_chain(x, layers, B, d, wscale, bscale, in_mode, epi):
    """layers: list of (W, bias or None, aux or None, out or None) -> one cluster launch (csrc/mlp.cu)."""
    global _mlp_checked
    lib = _lib.load()
    if not _mlp_checked:
        if lib.wgs_mlp_layer_size() != ctypes.sizeof(MlpLayer):
            raise RuntimeError('wgs_mlp_layer layout mismatch between header and ctypes mirror')
        _mlp_checked = True
    arr = (MlpLayer * len(layers))()
    for q, (W, bias, aux, out) in zip(arr, layers):
        for t in (W, bias, aux, out):
            if t is not None and not (t.is_cuda and t.is_contiguous()):
                raise RuntimeError('mapping network needs contiguous CUDA tensors (no CPU fallback)')
        q.W = W.data_ptr()
        q.bias = bias.data_ptr() if bias is not None else None
        q.aux = aux.data_ptr() if aux is not None else None
        q.out = out.data_ptr() if out is not None else None
    _lib.check(lib.wgs_mlp_chain(ctypes.c_void_p(x.data_ptr()), x.stride(0), arr, len(layers), B, d, float(wscale),
                                 float(bscale), int(in_mode), int(epi), _lib.stream()))


def _chain_ok(G, B):
    return MLP_CHAIN and G.style_dim % 128 == 0 and 2 * B * G.style_dim * 4 <= 200 * 1024 and 1 <= G.n_mlp <= 16


def _mapping_forward(G, z):
    """Returns the list [pixelnorm(z), h1, ..., h8 = w] (all kept for the backward pass)."""
    P = G.plan()
    B, d = z.shape
    acts = [torch.empty_like(z)]
    _lib.call('wgs_pixelnorm_rows', _lib.ptr(z), _lib.ptr(acts[0]), B, d, _lib.stream())
    wscale = (1.0 / math.sqrt(d)) * G.lr_mlp
    if _chain_ok(G, B) and d == G.style_dim:
        outs = torch.empty(G.n_mlp, B, G.style_dim, device=z.device, dtype=torch.float32)
        _mlp_chain(acts[0], [(P['map_w'][i], P['map_b'][i], None, outs[i]) for i in range(G.n_mlp)], B, d, wscale, G.lr_mlp, 0, 1)
        return acts + [outs[i] for i in range(G.n_mlp)]
    for i in range(G.n_mlp):
        out = torch.empty(B, G.style_dim, device=z.device, dtype=torch.float32)
        _linear(acts[-1], P['map_w'][i], P['map_b'][i], out, wscale=wscale, bscale=G.lr_mlp, epi=1)
        acts.append(out)
    return acts


def styles_and_demod(G, w):
    """All 26 per-layer styles in one launch, then the 17 demodulation vectors in one grouped launch."""
    P = G.plan()
    B = w.shape[0]
    s_all = torch.empty(B, P['sum_c'], device=w.device, dtype=torch.float32)
    _linear(w, P['mod_w'], P['mod_b'], s_all, wscale=1.0 / math.sqrt(G.style_dim), bscale=1.0)
    d_all = torch.empty(B * P['d_total'], device=w.device, dtype=torch.float32)
    demod, problems, off = [], [], 0
    for e in P['styled']:
        d = d_all[off: off + B * e['co']].view(B, e['co'])
        off += B * e['co']
        s = s_all[:, e['s_off']: e['s_off'] + e['ci']]
        problems.append(_problem(s, e['wsq'], d, wscale=e['scale'] ** 2, in_mode=1, epi=2, eps=1e-8))
        demod.append(d)
    _linear_group(problems, B)
    return s_all, demod


def synthesis(G, w, tape=None, grad_from=0):
    """w [B, style_dim] -> image NHWC [B, size, size, 3].  When `tape` is a dict, everything the
    data-gradient pass needs is recorded in it — for batch rows >= grad_from only (the un-shifted half of a
    paired batch needs no gradient: lib/trainer.py:200 feeds G(z) with a z that has no grad).

    Per layer: ONE tensor-core launch (4 for an up-sampling layer + the FIR kernel) whose epilogue writes the
    next layer's modulated split32 operand, accumulates ToRGB, and stores the fp32 activation only for the rows
    that will be back-propagated."""
    P = G.plan()
    B = w.shape[0]
    dev = w.device
    st = _lib.stream
    s_all, demod = styles_and_demod(G, w)
    g0 = grad_from
    if tape is not None:
        tape.update(s_all=s_all[g0:], demod=[d[g0:] for d in demod], acts=[], w=w[g0:])
    layers = P['styled']

    def style_of(e):
        return s_all[:, e['s_off']: e['s_off'] + e['ci']]

    a = P['const'].unsqueeze(0).expand(B, -1, -1, -1).contiguous()             # [B,4,4,C]
    xs = C.pack_split32(a, scale=style_of(layers[0]), rows_per_group=16)
    if tape is not None:
        tape['acts'].append(a[g0:])
    skip = None
    for li, e in enumerate(layers):
        n, h, wd = xs.shape[0], xs.shape[1], xs.shape[2]
        nxt = layers[li + 1] if li + 1 < len(layers) else None
        noise = P['noise'][li]
        oh, ow = (2 * h, 2 * wd) if e['up'] else (h, wd)
        a = torch.empty(n, oh, ow, e['co'], device=dev, dtype=torch.float32) if tape is not None else None
        xs_next = torch.empty(n, oh, ow, e['co'] // 32, 64, device=dev, dtype=torch.bfloat16) if nxt else None
        s_next = style_of(nxt) if nxt else None
        if e['up']:
            if 'w_up' in e:
                y = C.conv_transpose2d_s2_merged(xs, e['w_up'], 3, e['co'], split_k=2)    # [B, 2h+1, 2w+1, Co] raw
            else:
                y = C.conv_transpose2d_s2(xs, e['w_fwd'], 3, split_k=2)
            _lib.call('wgs_fir4_act', _lib.ptr(y), _lib.ptr(a), n, 2 * h + 1, 2 * wd + 1, oh, ow, e['co'], 1,
                      _TAPS, _lib.ptr(demod[li]), _lib.ptr(e['bias']), _lib.ptr(noise), e['noise_w'], 3,
                      _lib.ptr(xs_next), ctypes.c_void_p(s_next.data_ptr()), s_next.stride(0), g0, st())
        else:
            r = P['rgb'][li // 2]
            s_r = style_of(r)
            rgb = torch.empty(n, oh, ow, 3, device=dev, dtype=torch.float32)
            _lib.call('wgs_sg2_rgb_init', _lib.ptr(r['bias']), _lib.ptr(skip), _lib.ptr(rgb), n, oh, ow, _TAPS, st())
            wm = torch.empty(n, 3, e['co'], device=dev, dtype=torch.float32)
            _lib.call('wgs_sg2_rgb_weights', _lib.ptr(r['w']), ctypes.c_void_p(s_r.data_ptr()), s_r.stride(0),
                      _lib.ptr(wm), n, e['co'], r['scale'], st())
            C.conv2d(xs, e['w_fwd'], 3, 3, padding=1, out=a, no_f32=a is None, alpha=demod[li], beta=e['bias'],
                     noise=noise, noise_w=e['noise_w'], act=3, out_split=xs_next, split_scale=s_next, out_from_n=g0,
                     rgb_w=wm, rgb_out=rgb, split_k=2)
            skip = rgb
        if tape is not None:
            tape['acts'].append(a[g0:])
        xs = xs_next
    return skip


# ----------------------------------------------------------------------------------------------
# backward (data gradient only)
def synthesis_backward(G, tape, dimg):
    """dimg [B, size, size, 3] NHWC -> d(loss)/dw [B, style_dim] using the activations recorded in `tape`.

    Per layer (top to bottom): one fused boundary kernel (incoming conv gradient * style + ToRGB branch -> dpre,
    with the three style / demod reductions), [transposed FIR for up-sampling layers], one tensor-core
    data-gradient conv."""
    P = G.plan()
    s_all, demod, acts = tape['s_all'], tape['demod'], tape['acts']
    B = s_all.shape[0]
    dev = s_all.device
    ds_all = torch.zeros_like(s_all)
    drgb = dimg.contiguous()
    st = _lib.stream
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def sl(t, e):
        return t[:, e['s_off']: e['s_off'] + e['ci']]

    dd_all = torch.zeros(B * P['d_total'], device=dev, dtype=torch.float32)
    dd_off, demod_bwd = 0, []
    dx_up, up_e = None, None                     # gradient w.r.t. the modulated input of the layer above, and that layer
    for li in range(len(P['styled']) - 1, -1, -1):
        e = P['styled'][li]
        a = acts[li + 1]
        n, h, wd, co = a.shape
        npix = h * wd
        has_rgb = not e['up']
        r = P['rgb'][li // 2] if has_rgb else None
        dd = dd_all[dd_off: dd_off + n * co].view(n, co)
        dd_off += n * co
        if e['up']:
            dpre = torch.empty_like(a)
            gs = None
        else:
            dpre = None
            gs = torch.empty(n, h, wd, co // 32, 64, device=dev, dtype=torch.bfloat16)
        s_u = sl(s_all, up_e) if dx_up is not None else None
        ds_u = sl(ds_all, up_e) if dx_up is not None else None
        s_r = sl(s_all, r) if has_rgb else None
        ds_r = sl(ds_all, r) if has_rgb else None
        _lib.call('wgs_sg2_layer_bwd',
                  _lib.ptr(dx_up), vp(s_u) if s_u is not None else None, s_u.stride(0) if s_u is not None else 0,
                  vp(ds_u) if ds_u is not None else None, ds_u.stride(0) if ds_u is not None else 0,
                  _lib.ptr(drgb) if has_rgb else None, vp(s_r) if has_rgb else None, s_r.stride(0) if has_rgb else 0,
                  _lib.ptr(r['w']) if has_rgb else None, r['scale'] if has_rgb else 0.0,
                  vp(ds_r) if has_rgb else None, ds_r.stride(0) if has_rgb else 0,
                  _lib.ptr(a), _lib.ptr(demod[li]), _lib.ptr(e['bias']), _lib.ptr(P['noise'][li]), e['noise_w'],
                  _lib.ptr(dd), _lib.ptr(dpre), _lib.ptr(gs), n, npix, co, st())
        if has_rgb and li > 0:
            dprev = torch.empty(n, h // 2, wd // 2, 3, device=dev, dtype=torch.float32)
            _lib.call('wgs_sg2_rgb_up_bwd', _lib.ptr(drgb), _lib.ptr(dprev), n, h, wd, _TAPS, st())
            drgb = dprev
        if e['up']:
            # transposed blur * demod written straight as the split32 operand of the strided data-gradient conv
            hi, wi = h // 2, wd // 2
            gs = torch.empty(n, h + 1, wd + 1, co // 32, 64, device=dev, dtype=torch.bfloat16)
            _lib.call('wgs_fir4_act', _lib.ptr(dpre), None, n, h, wd, h + 1, wd + 1, co, 2, _TAPS,
                      _lib.ptr(demod[li]), None, None, 0.0, 0, _lib.ptr(gs), None, 0, 0, st())
            dx = C.conv2d(gs, e['w_bwd'], 3, 3, stride=2, padding=0, cout=e['ci'], split_k=2)       # [n, hi, wi, ci]
            assert dx.shape[1] == hi and dx.shape[2] == wi
        else:
            dx = C.conv2d(gs, e['w_bwd'], 3, 3, padding=1, cout=e['ci'], split_k=2)
        # demodulation: d = rsqrt(scale^2 sum_i s_i^2 Wsq[o,i] + eps)  ->  ds_i += s_i * sum_o (-dd_o d_o^3 scale^2) Wsq[o,i]
        # (deferred: all layers in one grouped launch after the loop; every writer of ds_all accumulates)
        s_e, ds_e = sl(s_all, e), sl(ds_all, e)
        demod_bwd.append(_problem(dd, e['wsq_t'], ds_e, x2=demod[li], mul=s_e, wscale=-(e['scale'] ** 2), in_mode=3,
                                  accumulate=1))
        if li == 0:
            pin = dx.shape[1] * dx.shape[2]
            _lib.call('wgs_sg2_mod_bwd', _lib.ptr(dx), _lib.ptr(P['const']), 1, vp(s_e), s_e.stride(0), None, 0,
                      vp(ds_e), ds_e.stride(0), n, pin, e['ci'], st())
        dx_up, up_e = dx, e
    _linear_group(demod_bwd, B)
    # dw = ds_all @ mod_w (512 outputs over ~9000 style channels): split along the contraction into chunks that atomically
    # share the output - one warp per output streaming a 35 KB weight row left this launch at 128 CTAs / 122 us
    dw = torch.zeros(B, G.style_dim, device=dev, dtype=torch.float32)
    sum_c, step = P['sum_c'], 1024
    parts = [_problem(ds_all[:, lo: min(sum_c, lo + step)], P['mod_w_t'][:, lo: min(sum_c, lo + step)], dw,
                      wscale=1.0 / math.sqrt(G.style_dim), accumulate=2) for lo in range(0, sum_c, step)]
    _linear_group(parts, B)
    return dw


def _mapping_backward(G, acts, dw):
    """Backward of PixelNorm + n_mlp fused-lrelu EqualLinears; acts = [pixelnorm(z), h1..hn]."""
    P = G.plan()
    wscale = (1.0 / math.sqrt(G.style_dim)) * G.lr_mlp
    g = dw
    B = g.shape[0]
    if _chain_ok(G, B) and all(a.is_contiguous() for a in acts):
        out = torch.empty_like(acts[0])
        n = G.n_mlp
        # the fused-lrelu derivative (taken from each layer's forward output) is applied to g inside the chain
        layers = [(P['map_w_t'][i], None, acts[i + 1], out if i == 0 else None) for i in range(n - 1, -1, -1)]
        _mlp_chain(g.contiguous(), layers, B, G.style_dim, wscale, 0.0, 2, 0)
        return out
    for i in range(G.n_mlp - 1, -1, -1):
        out = torch.empty_like(acts[i])
        # the fused-lrelu derivative (taken from the layer's forward output) is applied to g while it is loaded
        _linear_group([_problem(g, P['map_w_t'][i], out, x2=acts[i + 1], wscale=wscale, in_mode=2)], g.shape[0])
        g = out
    return g


class _MappingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, G, z):
        acts = _mapping_forward(G, z)
        ctx.G = G
        ctx.acts = acts
        ctx.save_for_backward(z)
        return acts[-1]

    @staticmethod
    def backward(ctx, dw):
        (z,) = ctx.saved_tensors
        g = _mapping_backward(ctx.G, ctx.acts, dw.contiguous())
        # PixelNorm: y = z * r, r = rsqrt(mean(z^2) + 1e-8)  ->  dz = r*g - z * r^3 * mean(z*g)
        r = torch.rsqrt(z.pow(2).mean(dim=1, keepdim=True) + 1e-8)
        dz = r * g - z * r.pow(3) * (z * g).mean(dim=1, keepdim=True)
        return None, dz


class _SynthesisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, G, w):
        need = ctx.needs_input_grad[1]
        tape = {} if need else None
        img = synthesis(G, w.detach().contiguous(), tape)
        ctx.G, ctx.tape = G, tape
        return img

    @staticmethod
    def backward(ctx, dimg):
        if ctx.tape is None:
            return None, None
        dw = synthesis_backward(ctx.G, ctx.tape, dimg)
        ctx.tape = None
        return None, dw


class _PairFn(torch.autograd.Function):
    """Both halves of the batched pass as separate outputs: the gradient of the shifted half arrives as it is (no
    zero-padded full-batch gradient, no slice copies)."""

    @staticmethod
    def forward(ctx, G, w_all, n_plain):
        need = ctx.needs_input_grad[1]
        tape = {} if need else None
        img = synthesis(G, w_all.detach(), tape, grad_from=n_plain)
        ctx.G, ctx.tape, ctx.n_plain, ctx.rows = G, tape, n_plain, w_all.shape[0]
        plain, shifted = img[:n_plain], img[n_plain:]
        ctx.mark_non_differentiable(plain)
        return plain, shifted

    @staticmethod
    def backward(ctx, _dplain, dimg):
        if ctx.tape is None or dimg is None:
            return None, None, None
        dw = synthesis_backward(ctx.G, ctx.tape, dimg)
        ctx.tape = None
        full = dw.new_zeros(ctx.rows, dw.shape[1])
        full[ctx.n_plain:] = dw
        return None, full, None
