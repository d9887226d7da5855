"""Oracle: ProgGAN generator (test infrastructure, CPU torch, functional over a state dict).

Keys are the reference module's: ``features.{i}.conv.weight``, ``features.{i}.wscale.scale``,
``features.{i}.wscale.b``, ``output.conv.weight``, ``output.wscale.*``.  Follows
/root/reference/models/ProgGAN/model.py:12-95 and ProgGANWrapper (models/gan_load.py:109-120).
"""
import math
import torch
import torch.nn.functional as F

# (in, out, kernel, padding, upscale) for features.0 .. features.17   (model.py:68-86)
PLAN_1024 = (
    [(512, 512, 4, 3, False), (512, 512, 3, 1, False)]
    + [(512, 512, 3, 1, True), (512, 512, 3, 1, False)] * 3
    + [(512, 256, 3, 1, True), (256, 256, 3, 1, False),
       (256, 128, 3, 1, True), (128, 128, 3, 1, False),
       (128, 64, 3, 1, True), (64, 64, 3, 1, False),
       (64, 32, 3, 1, True), (32, 32, 3, 1, False),
       (32, 16, 3, 1, True), (16, 16, 3, 1, False)]
)


def pixel_norm(x):
    """x / sqrt(mean_c x^2 + 1e-8)  (model.py:17-18)."""
    return x / torch.sqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


def generate(sd, z, shift=None, plan=PLAN_1024):
    """ProgGANWrapper.forward: reshape z(+shift) to [B, 512, 1, 1], 18 norm-(up)-conv-wscale-lrelu
    blocks, then norm - 1x1 conv - wscale (no activation, no clamp)."""
    x = (z if shift is None else z + shift)
    x = x.reshape(x.shape[0], x.shape[1], 1, 1)
    for i, (_, _, k, pad, up) in enumerate(plan):
        x = pixel_norm(x)
        if up:
            x = F.interpolate(x, scale_factor=2, mode='nearest')
        x = F.conv2d(x, sd['features.%d.conv.weight' % i], None, 1, pad)
        x = x * sd['features.%d.wscale.scale' % i] + sd['features.%d.wscale.b' % i].view(1, -1, 1, 1)
        x = F.leaky_relu(x, 0.2)
    x = pixel_norm(x)
    x = F.conv2d(x, sd['output.conv.weight'])
    return x * sd['output.wscale.scale'] + sd['output.wscale.b'].view(1, -1, 1, 1)


def init_state(plan=PLAN_1024, generator=None, pretrained_like=True):
    """Random init.  ``pretrained_like=True`` uses conv.weight ~ N(0,1), scale = sqrt(2/fan_in), b = 0 —
    the statistics of the released model — because the reference constructor's own init
    (scale, b ~ randn; default-init convs) gives bias-dominated outputs that hide conv errors
    (SURVEY.md §7 "Random-init weights can hide errors").  ``False`` reproduces the constructor."""
    g = generator
    sd = {}
    convs = [('features.%d' % i, ci, co, k) for i, (ci, co, k, _, _) in enumerate(plan)]
    convs.append(('output', plan[-1][1], 3, 1))
    for name, ci, co, k in convs:
        fan_in = ci * k * k
        if pretrained_like:
            sd[name + '.conv.weight'] = torch.randn(co, ci, k, k, generator=g)
            gain = 1.0 if name == 'output' else math.sqrt(2.0)
            sd[name + '.wscale.scale'] = torch.tensor([gain / math.sqrt(fan_in)])
            sd[name + '.wscale.b'] = torch.zeros(co)
        else:
            bound = 1.0 / math.sqrt(fan_in)
            sd[name + '.conv.weight'] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
            sd[name + '.wscale.scale'] = torch.randn(1, generator=g)
            sd[name + '.wscale.b'] = torch.randn(co, generator=g)
    return sd
