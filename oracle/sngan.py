"""Oracle: SNGAN ResNet generator in eval mode (test infrastructure, CPU torch, functional).

Keys are those of the wrapped ``GenWrapper.model`` Sequential
(/root/reference/models/SNGAN/sn_gen_resnet.py:81-112):
  ``0.{weight,bias}``                       dense 128 -> 4*4*C0
  ``{2+i}.conv1.*``, ``{2+i}.conv2.*``      residual block convs (also visible as ``model.3/6``)
  ``{2+i}.model.{0,4}.*``                   the two BatchNorms of block i
  ``{2+i}.bypass.1.*``                      3x3 bypass conv when channels change
  ``{n+2}.*`` BatchNorm, ``{n+4}.*`` final conv
Follows ResBlockGenerator (:24-54), make_resnet_generator (:81-112) and SNGANWrapper
(models/gan_load.py:21-28).  BatchNorm uses running statistics (the trainer puts G in eval mode,
lib/trainer.py:144).
"""
import math
import torch
import torch.nn.functional as F

CONFIGS = {
    'sn_resnet32': ([256, 256, 256, 256], 4),
    'sn_resnet64': ([1024, 512, 256, 128, 64], 4),
}


def _bn_eval(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'],
                        False, 0.0, eps)


def _up(x):
    return F.interpolate(x, scale_factor=2, mode='nearest')


def generate(sd, z, shift=None, model='sn_resnet32'):
    channels, seed = CONFIGS[model]
    x = z if shift is None else z + shift
    x = F.linear(x, sd['0.weight'], sd['0.bias']).view(-1, channels[0], seed, seed)
    for i in range(len(channels) - 1):
        p = '%d' % (2 + i)
        h = F.relu(_bn_eval(sd, p + '.model.0', x))
        h = F.conv2d(_up(h), sd[p + '.conv1.weight'], sd[p + '.conv1.bias'], padding=1)
        h = F.relu(_bn_eval(sd, p + '.model.4', h))
        h = F.conv2d(h, sd[p + '.conv2.weight'], sd[p + '.conv2.bias'], padding=1)
        s = _up(x)
        if channels[i] != channels[i + 1]:
            s = F.conv2d(s, sd[p + '.bypass.1.weight'], sd[p + '.bypass.1.bias'], padding=1)
        x = h + s
    n = len(channels) + 1
    x = F.relu(_bn_eval(sd, '%d' % n, x))
    x = F.conv2d(x, sd['%d.weight' % (n + 2)], sd['%d.bias' % (n + 2)], padding=1)
    return torch.tanh(x)


def _xavier(shape, gain, g):
    fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
    fan_out = shape[0] * (shape[2] * shape[3] if len(shape) == 4 else 1)
    a = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(*shape, generator=g) * 2 - 1) * a


def init_state(model='sn_resnet32', image_channels=1, latent_dim=128, generator=None):
    """Random init following the constructor (xavier-uniform convs, default BN); running
    statistics are perturbed away from (0, 1) so that eval-mode BN is not a no-op."""
    g = generator
    channels, seed = CONFIGS[model]
    sd = {'0.weight': _xavier((seed * seed * channels[0], latent_dim), 1.0, g),
          '0.bias': (torch.rand(seed * seed * channels[0], generator=g) * 2 - 1) / math.sqrt(latent_dim)}

    def bn(p, c):
        sd[p + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[p + '.bias'] = 0.1 * torch.randn(c, generator=g)
        sd[p + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
        sd[p + '.running_var'] = 1.0 + 0.2 * torch.rand(c, generator=g)

    def conv(p, ci, co, gain):
        sd[p + '.weight'] = _xavier((co, ci, 3, 3), gain, g)
        sd[p + '.bias'] = (torch.rand(co, generator=g) * 2 - 1) / math.sqrt(ci * 9)

    for i in range(len(channels) - 1):
        p = '%d' % (2 + i)
        ci, co = channels[i], channels[i + 1]
        conv(p + '.conv1', ci, co, math.sqrt(2))
        conv(p + '.conv2', co, co, math.sqrt(2))
        bn(p + '.model.0', ci)
        bn(p + '.model.4', co)
        if ci != co:
            conv(p + '.bypass.1', ci, co, 1.0)
    n = len(channels) + 1
    bn('%d' % n, channels[-1])
    conv('%d' % (n + 2), channels[-1], image_channels, 1.0)
    return sd
