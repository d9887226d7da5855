"""ProgGAN, SNGAN and BigGAN generators with every convolution on the tcgen05 tap-list kernel.

State-dict keys equal the reference modules' (models/ProgGAN/model.py:65-95, models/SNGAN/sn_gen_resnet.py:81-112,
models/BigGAN/BigGAN.py:54-243), so released checkpoints load with ``load_state_dict``.  The generators are
frozen during WarpedGANSpace training: weights are folded once per ``plan()`` (ProgGAN WScale into the conv
weights; BigGAN spectral norm W/sigma — the reference redoes a non-updating power iteration on every forward,
models/BigGAN/layers.py:84-96; eval-mode BatchNorm into a per-channel affine) and only the data gradient is
propagated.  Round-1 state: the convolutions (forward + data-gradient) are ours; the light glue between them
(pixel norm, nearest upsample, ReLU/tanh, attention softmax/bmm) is still ATen and is the next fusion target.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from . import _tree
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
    """models/ProgGAN/model.py:65-95."""

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
        if self._plan is None:
            t = _tree.tensors(self)
            with torch.no_grad():
                self._plan = {n: (t[n + '.conv.weight'] * t[n + '.wscale.scale']).detach().contiguous()
                              for n in ['features.%d' % i for i in range(len(self.blocks))] + ['output']}
        return self._plan

    def forward(self, x):
        self._require_cuda(x)
        t, w = _tree.tensors(self), self.plan()
        for i, (_, _, k, pad, up) in enumerate(self.blocks):
            x = _pixel_norm(x)
            if up:
                x = F.interpolate(x, scale_factor=2, mode='nearest')
            n = 'features.%d' % i
            x = F.leaky_relu(conv2d(_cl(x), w[n], t[n + '.wscale.b'].detach(), 1, pad), 0.2)
        return conv2d(_cl(_pixel_norm(x)), w['output'], t['output.wscale.b'].detach(), 1, 0)


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
        """W / sigma for every spectrally-normalised layer (one power iteration off the stored u0, as the
        reference's eval forward does every call)."""
        if self._plan is None:
            t = {k: v.detach() for k, v in _tree.tensors(self).items()}
            out = {}
            with torch.no_grad():
                for k in t:
                    if k.endswith('.u0'):
                        p = k[:-3]
                        w = t[p + '.weight']
                        wm = w.reshape(w.shape[0], -1)
                        v = F.normalize(t[k] @ wm, eps=self.SN_eps)
                        u2 = F.normalize(v @ wm.t(), eps=self.SN_eps)
                        out[p] = (w / torch.squeeze((v @ wm.t()) @ u2.t())).contiguous()
            self._plan = out
        return self._plan

    def _ccbn(self, t, w, p, x, y):
        gain = 1.0 + F.linear(y, w[p + '.gain'])
        bias = F.linear(y, w[p + '.bias'])
        inv = torch.rsqrt(t[p + '.stored_var'] + self.BN_eps)
        xn = (x - t[p + '.stored_mean'].view(1, -1, 1, 1)) * inv.view(1, -1, 1, 1)
        return xn * gain.view(x.shape[0], -1, 1, 1) + bias.view(x.shape[0], -1, 1, 1)

    def forward(self, z, y):
        self._require_cuda(z)
        t = {k: v.detach() for k, v in _tree.tensors(self).items()}
        w = self.plan()
        nb = len(self.arch['out'])
        if self.hier:
            zs = torch.split(z, self.z_chunk_size, 1)
            z = zs[0]
            ys = [torch.cat([y, item], 1) for item in zs[1:]]
        else:
            ys = [y] * nb
        h = F.linear(z, w['linear'], t['linear.bias']).view(z.shape[0], -1, self.bottom_width, self.bottom_width)
        for i in range(nb):
            p = 'blocks.%d.0' % i
            x = h
            h = F.interpolate(F.relu(self._ccbn(t, w, p + '.bn1', x, ys[i])), scale_factor=2)
            x = F.interpolate(x, scale_factor=2)
            h = conv2d(_cl(h), w[p + '.conv1'], t[p + '.conv1.bias'], 1, 1)
            h = conv2d(_cl(F.relu(self._ccbn(t, w, p + '.bn2', h, ys[i]))), w[p + '.conv2'], t[p + '.conv2.bias'], 1, 1)
            h = h + conv2d(_cl(x), w[p + '.conv_sc'], t[p + '.conv_sc.bias'], 1, 0)
            if self.arch['attn'][i]:
                q = 'blocks.%d.1' % i
                b, c, hh, ww = h.shape
                hc = _cl(h)
                theta = conv2d(hc, w[q + '.theta'], None, 1, 0).reshape(b, c // 8, hh * ww)
                phi = F.max_pool2d(conv2d(hc, w[q + '.phi'], None, 1, 0), [2, 2]).reshape(b, c // 8, hh * ww // 4)
                g = F.max_pool2d(conv2d(hc, w[q + '.g'], None, 1, 0), [2, 2]).reshape(b, c // 2, hh * ww // 4)
                beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
                o = torch.bmm(g, beta.transpose(1, 2)).view(b, c // 2, hh, ww)
                h = t[q + '.gamma'] * conv2d(_cl(o), w[q + '.o'], None, 1, 0) + h
        p = 'output_layer.0'
        inv = t[p + '.gain'] * torch.rsqrt(t[p + '.stored_var'] + self.BN_eps)
        h = F.relu((h - t[p + '.stored_mean'].view(1, -1, 1, 1)) * inv.view(1, -1, 1, 1) + t[p + '.bias'].view(1, -1, 1, 1))
        return torch.tanh(conv2d(_cl(h), w['output_layer.2'], t['output_layer.2.bias'], 1, 1))
