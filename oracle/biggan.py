"""Oracle: BigGAN generator in eval mode (test infrastructure, CPU torch, functional).

Keys are the reference ``BigGAN.Generator`` state dict's (``shared.weight``, ``linear.*``,
``blocks.{i}.0.{conv1,conv2,conv_sc}.*``, ``blocks.{i}.0.bn{1,2}.{gain,bias}.weight``,
``blocks.{i}.0.bn{1,2}.stored_{mean,var}``, ``blocks.{i}.1.{theta,phi,g,o}.weight``, ``blocks.{i}.1.gamma``,
``output_layer.0.*``, ``output_layer.2.*``, plus the ``u0`` power-iteration buffer of every SN layer).
Follows /root/reference/models/BigGAN/BigGAN.py:13-52 (arch), :222-243 (forward);
layers.py:25-47,84-96 (spectral norm: one non-updating power iteration per forward in eval),
:141-166 (attention), :275-322 (ccbn), :330-363 (bn), :372-405 (GBlock);
BigGANWrapper (models/gan_load.py:65-81).
"""
import torch
import torch.nn.functional as F


def arch(resolution, ch=96, attention='64'):
    """BigGAN.py:13-52 for the resolutions the reference can instantiate."""
    mult = {
        256: ([16, 16, 8, 8, 4, 2], [16, 8, 8, 4, 2, 1]),
        128: ([16, 16, 8, 4, 2], [16, 8, 4, 2, 1]),
        64: ([16, 16, 8, 4], [16, 8, 4, 2]),
        32: ([4, 4, 4], [4, 4, 4]),
    }[resolution]
    attn = [int(a) for a in attention.split('_')]
    res = [8 * 2 ** i for i in range(len(mult[0]))]
    return {'in': [ch * m for m in mult[0]], 'out': [ch * m for m in mult[1]], 'res': res,
            'attn': [r in attn for r in res]}


def sn_weight(sd, p, eps=1e-6):
    """W / sigma with sigma from one power iteration off the stored u0 (layers.py:25-47,84-96)."""
    w = sd[p + '.weight']
    wm = w.reshape(w.shape[0], -1)
    u = sd[p + '.u0']
    v = F.normalize(u @ wm, eps=eps)
    u2 = F.normalize(v @ wm.t(), eps=eps)
    sigma = torch.squeeze((v @ wm.t()) @ u2.t())
    return w / sigma


def ccbn(sd, p, x, y, eps=1e-5, sn_eps=1e-6):
    """Class-conditional BN, eval mode (layers.py:303-322): stored stats, then *(1+gain(y)) + bias(y)."""
    gain = 1.0 + F.linear(y, sn_weight(sd, p + '.gain', sn_eps))
    bias = F.linear(y, sn_weight(sd, p + '.bias', sn_eps))
    out = F.batch_norm(x, sd[p + '.stored_mean'], sd[p + '.stored_var'], None, None, False, 0.1, eps)
    return out * gain.view(x.shape[0], -1, 1, 1) + bias.view(x.shape[0], -1, 1, 1)


def _conv(sd, p, x, pad, sn_eps):
    return F.conv2d(x, sn_weight(sd, p, sn_eps), sd.get(p + '.bias'), padding=pad)


def gblock(sd, p, x, y, sn_eps=1e-6):
    """layers.py:395-405 with upsample = nearest x2."""
    h = F.relu(ccbn(sd, p + '.bn1', x, y, sn_eps=sn_eps))
    h = F.interpolate(h, scale_factor=2)
    x = F.interpolate(x, scale_factor=2)
    h = _conv(sd, p + '.conv1', h, 1, sn_eps)
    h = F.relu(ccbn(sd, p + '.bn2', h, y, sn_eps=sn_eps))
    h = _conv(sd, p + '.conv2', h, 1, sn_eps)
    return h + _conv(sd, p + '.conv_sc', x, 0, sn_eps)


def attention(sd, p, x, sn_eps=1e-6):
    """layers.py:153-166."""
    b, ch, hh, ww = x.shape
    theta = _conv(sd, p + '.theta', x, 0, sn_eps).view(b, ch // 8, hh * ww)
    phi = F.max_pool2d(_conv(sd, p + '.phi', x, 0, sn_eps), [2, 2]).view(b, ch // 8, hh * ww // 4)
    g = F.max_pool2d(_conv(sd, p + '.g', x, 0, sn_eps), [2, 2]).view(b, ch // 2, hh * ww // 4)
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = _conv(sd, p + '.o', torch.bmm(g, beta.transpose(1, 2)).view(b, ch // 2, hh, ww), 0, sn_eps)
    return sd[p + '.gamma'] * o + x


def generate(sd, z, classes, shift=None, resolution=128, ch=96, attention_at='64', bottom_width=4,
             hier=True, sn_eps=1e-6, bn_eps=1e-5):
    """BigGANWrapper.forward with explicit class ids: G(z(+shift), shared(classes))."""
    a = arch(resolution, ch, attention_at)
    x = z if shift is None else z + shift
    y = F.embedding(classes, sd['shared.weight'])
    nblocks = len(a['out'])
    if hier:
        chunk = x.shape[1] // (nblocks + 1)
        zs = torch.split(x, chunk, 1)
        x = zs[0]
        ys = [torch.cat([y, zc], 1) for zc in zs[1:]]
    else:
        ys = [y] * nblocks
    h = F.linear(x, sn_weight(sd, 'linear', sn_eps), sd['linear.bias'])
    h = h.view(h.shape[0], -1, bottom_width, bottom_width)
    for i in range(nblocks):
        h = gblock(sd, 'blocks.%d.0' % i, h, ys[i], sn_eps)
        if a['attn'][i]:
            h = attention(sd, 'blocks.%d.1' % i, h, sn_eps)
    p = 'output_layer.0'
    h = F.batch_norm(h, sd[p + '.stored_mean'], sd[p + '.stored_var'], sd[p + '.gain'], sd[p + '.bias'],
                     False, 0.1, bn_eps)
    h = _conv(sd, 'output_layer.2', F.relu(h), 1, sn_eps)
    return torch.tanh(h)


def init_state(resolution=128, ch=96, dim_z=120, shared_dim=128, n_classes=1000, attention_at='64',
               bottom_width=4, generator=None):
    """Random init: N(0, 0.02)-ish weights scaled to unit spectral norm by SN anyway; stored BN stats
    perturbed away from (0,1); attention gamma set to 0.5 so the attention path contributes."""
    g = generator
    a = arch(resolution, ch, attention_at)
    nblocks = len(a['out'])
    chunk = dim_z // (nblocks + 1)
    cond = shared_dim + chunk
    sd = {'shared.weight': torch.randn(n_classes, shared_dim, generator=g)}

    def sn(p, shape, bias=True):
        sd[p + '.weight'] = torch.randn(*shape, generator=g) * 0.05
        if bias:
            sd[p + '.bias'] = torch.randn(shape[0], generator=g) * 0.05
        sd[p + '.u0'] = torch.randn(1, shape[0], generator=g)
        sd[p + '.sv0'] = torch.ones(1)

    def stats(p, c):
        sd[p + '.stored_mean'] = 0.1 * torch.randn(c, generator=g)
        sd[p + '.stored_var'] = 1.0 + 0.2 * torch.rand(c, generator=g)

    sn('linear', (a['in'][0] * bottom_width ** 2, chunk))
    for i in range(nblocks):
        p = 'blocks.%d.0' % i
        ci, co = a['in'][i], a['out'][i]
        sn(p + '.conv1', (co, ci, 3, 3))
        sn(p + '.conv2', (co, co, 3, 3))
        sn(p + '.conv_sc', (co, ci, 1, 1))
        for name, c in (('.bn1', ci), ('.bn2', co)):
            sn(p + name + '.gain', (c, cond), bias=False)
            sn(p + name + '.bias', (c, cond), bias=False)
            stats(p + name, c)
        if a['attn'][i]:
            q = 'blocks.%d.1' % i
            sn(q + '.theta', (co // 8, co, 1, 1), bias=False)
            sn(q + '.phi', (co // 8, co, 1, 1), bias=False)
            sn(q + '.g', (co // 2, co, 1, 1), bias=False)
            sn(q + '.o', (co, co // 2, 1, 1), bias=False)
            sd[q + '.gamma'] = torch.tensor(0.5)
    sd['output_layer.0.gain'] = 1.0 + 0.1 * torch.randn(a['out'][-1], generator=g)
    sd['output_layer.0.bias'] = 0.1 * torch.randn(a['out'][-1], generator=g)
    stats('output_layer.0', a['out'][-1])
    sn('output_layer.2', (3, a['out'][-1], 3, 3))
    return sd
