"""Oracle: StyleGAN2 generator (test infrastructure, CPU torch, functional over a state dict).

State-dict keys are the reference module's (models/StyleGAN2/model.py:285-335), so a
reference ``Generator.state_dict()`` can be fed in directly.  Follows:
  * mapping net (PixelNorm + 8 EqualLinear, lr_mul 0.01, fused lrelu)  model.py:9-15,110-131,291-295
  * ModulatedConv2d (modulate, demodulate, grouped (transposed) conv)   model.py:150-228
  * Blur / Upsample / upfirdn2d                                        model.py:29-81, op/upfirdn2d.py:152-186
  * NoiseInjection + FusedLeakyReLU                                    model.py:231-241, op/fused_bias_act_kernel.cu:25-47
  * ToRGB with skip upsample                                           model.py:270-282
  * Generator.forward (fixed noise buffers, 18 identical style slots)   model.py:359-408
  * StyleGAN2Wrapper Z/W logic                                         models/gan_load.py:157-179
"""
import math
import torch
import torch.nn.functional as F

SQRT2 = 2.0 ** 0.5


def default_channels(channel_multiplier=2):
    """model.py:297-307."""
    c = {4: 512, 8: 512, 16: 512, 32: 512}
    for res, base in ((64, 256), (128, 128), (256, 64), (512, 32), (1024, 16)):
        c[res] = base * channel_multiplier
    return c


def fir_kernel(taps=(1, 3, 3, 1)):
    """make_kernel, model.py:18-26: separable outer product, normalised to sum 1."""
    k = torch.tensor(taps, dtype=torch.float32)
    k = torch.outer(k, k)
    return k / k.sum()


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """x [N,C,H,W]; zero-insert x up, pad (negative = crop), correlate with the
    flipped kernel, decimate x down (op/upfirdn2d.py:152-186 applied per channel)."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    p0, p1 = pad
    if up > 1:
        y = x.new_zeros(n, c, h, up, w, up)
        y[:, :, :, 0, :, 0] = x
        x = y.reshape(n, c, h * up, w * up)
    x = F.pad(x, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    x = x[:, :, max(-p0, 0): x.shape[2] - max(-p1, 0), max(-p0, 0): x.shape[3] - max(-p1, 0)]
    wgt = torch.flip(kernel, [0, 1]).to(device=x.device, dtype=x.dtype).view(1, 1, kh, kw).expand(c, 1, kh, kw)
    x = F.conv2d(x, wgt, groups=c)
    return x[:, :, ::down, ::down]


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=SQRT2):
    """scale * lrelu(x + b[channel]) (op/fused_bias_act_kernel.cu:25-47, act=3, grad=0)."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return scale * F.leaky_relu(x + bias.view(*shape), negative_slope)


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """model.py:110-131."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    y = F.linear(x, weight * scale)
    if activation:
        return fused_leaky_relu(y, bias * lr_mul)
    return y + bias * lr_mul


def mapping(sd, z, n_mlp=8, lr_mlp=0.01):
    """Generator.style: PixelNorm then n_mlp EqualLinear+fused lrelu (model.py:291-295)."""
    x = z * torch.rsqrt(torch.mean(z * z, dim=1, keepdim=True) + 1e-8)
    for i in range(1, n_mlp + 1):
        x = equal_linear(x, sd['style.%d.weight' % i], sd['style.%d.bias' % i], lr_mul=lr_mlp, activation=True)
    return x


def modulated_conv(sd, prefix, x, w, demodulate=True, upsample=False, blur_taps=(1, 3, 3, 1)):
    """ModulatedConv2d.forward (model.py:187-228), without the downsample branch
    (never built by the generator)."""
    weight = sd[prefix + '.weight']                              # [1, Co, Ci, k, k]
    _, co, ci, k, _ = weight.shape
    b, _, h, wd = x.shape
    style = equal_linear(w, sd[prefix + '.modulation.weight'], sd[prefix + '.modulation.bias'])
    wmod = (1.0 / math.sqrt(ci * k * k)) * weight * style.view(b, 1, ci, 1, 1)
    if demodulate:
        wmod = wmod * torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8).view(b, co, 1, 1, 1)
    if upsample:
        wt = wmod.transpose(1, 2).reshape(b * ci, co, k, k)
        y = F.conv_transpose2d(x.reshape(1, b * ci, h, wd), wt, padding=0, stride=2, groups=b)
        y = y.view(b, co, y.shape[2], y.shape[3])
        p = (len(blur_taps) - 2) - (k - 1)
        kern = fir_kernel(blur_taps) * 4.0
        return upfirdn2d(y, kern.to(y.dtype), pad=((p + 1) // 2 + 1, p // 2 + 1))
    y = F.conv2d(x.reshape(1, b * ci, h, wd), wmod.view(b * co, ci, k, k), padding=k // 2, groups=b)
    return y.view(b, co, y.shape[2], y.shape[3])


def styled_conv(sd, prefix, x, w, noise, upsample=False):
    """StyledConv.forward: conv -> + noise_weight * noise -> fused lrelu (model.py:253-267)."""
    y = modulated_conv(sd, prefix + '.conv', x, w, demodulate=True, upsample=upsample)
    y = y + sd[prefix + '.noise.weight'] * noise
    return fused_leaky_relu(y, sd[prefix + '.activate.bias'])


def to_rgb(sd, prefix, x, w, skip=None):
    """ToRGB.forward (model.py:270-282); skip upsample = upfirdn2d(up=2, 4*k, pad (2,1))."""
    y = modulated_conv(sd, prefix + '.conv', x, w, demodulate=False) + sd[prefix + '.bias']
    if skip is not None:
        kern = fir_kernel() * 4.0
        y = y + upfirdn2d(skip, kern.to(y.dtype), up=2, pad=(2, 1))
    return y


def synthesis(sd, w, size):
    """Generator.forward body after the mapping net, one w per sample broadcast to all
    n_latent slots, fixed noise buffers (model.py:364-405)."""
    log_size = int(math.log2(size))
    b = w.shape[0]
    x = sd['input.input'].repeat(b, 1, 1, 1)
    x = styled_conv(sd, 'conv1', x, w, sd['noises.noise_0'])
    skip = to_rgb(sd, 'to_rgb1', x, w)
    for i in range(log_size - 2):
        x = styled_conv(sd, 'convs.%d' % (2 * i), x, w, sd['noises.noise_%d' % (2 * i + 1)], upsample=True)
        x = styled_conv(sd, 'convs.%d' % (2 * i + 1), x, w, sd['noises.noise_%d' % (2 * i + 2)])
        skip = to_rgb(sd, 'to_rgbs.%d' % i, x, w, skip)
    return skip


def generate(sd, z, shift=None, size=1024, shift_in_w_space=False, latent_is_w=False):
    """StyleGAN2Wrapper.forward (models/gan_load.py:157-179)."""
    if shift_in_w_space:
        w = z if latent_is_w else mapping(sd, z)
        return synthesis(sd, w if shift is None else w + shift, size)
    return synthesis(sd, mapping(sd, z if shift is None else z + shift), size)


def init_state(size=1024, style_dim=512, n_mlp=8, channels=None, channel_multiplier=2,
               lr_mlp=0.01, generator=None, noise_strength=0.1):
    """Random-init state dict with the reference constructor's distributions
    (model.py:88,113,182-183,234,247,326) — except noise.weight, which the reference
    zero-initialises; it is set to ``noise_strength`` so the noise path is exercised
    (SURVEY.md §8d)."""
    ch = channels or default_channels(channel_multiplier)
    g = generator
    rn = lambda *s: torch.randn(*s, generator=g)
    sd = {}
    for i in range(1, n_mlp + 1):
        sd['style.%d.weight' % i] = rn(style_dim, style_dim) / lr_mlp
        sd['style.%d.bias' % i] = torch.zeros(style_dim)
    sd['input.input'] = rn(1, ch[4], 4, 4)

    def mod_conv(prefix, ci, co, k):
        sd[prefix + '.weight'] = rn(1, co, ci, k, k)
        sd[prefix + '.modulation.weight'] = rn(ci, style_dim)
        sd[prefix + '.modulation.bias'] = torch.ones(ci)

    def styled(prefix, ci, co):
        mod_conv(prefix + '.conv', ci, co, 3)
        sd[prefix + '.noise.weight'] = torch.full((1,), float(noise_strength))
        sd[prefix + '.activate.bias'] = torch.zeros(co)

    def rgb(prefix, ci):
        mod_conv(prefix + '.conv', ci, 3, 1)
        sd[prefix + '.bias'] = torch.zeros(1, 3, 1, 1)

    styled('conv1', ch[4], ch[4])
    rgb('to_rgb1', ch[4])
    log_size = int(math.log2(size))
    cin = ch[4]
    for i in range(log_size - 2):
        co = ch[2 ** (i + 3)]
        styled('convs.%d' % (2 * i), cin, co)
        styled('convs.%d' % (2 * i + 1), co, co)
        rgb('to_rgbs.%d' % i, co)
        cin = co
    for layer in range((log_size - 2) * 2 + 1):
        res = (layer + 5) // 2
        sd['noises.noise_%d' % layer] = rn(1, 1, 2 ** res, 2 ** res)
    return sd
