"""Oracle: Reconstructor (test infrastructure, CPU torch, functional over a state dict).

Keys are the reference module's (lib/reconstructor.py:10-69): ResNet — ``features_extractor.*``
(torchvision resnet18 names, conv1 swapped for a 6-channel 7x7/2 conv; the unused ``fc.*`` is kept in
the state dict but never read), ``path_indices.*``, ``shift_magnitudes.*``; LeNet —
``feature_extractor.{0,1,4,5,8,9}.*`` and heads ``path_indices.{0,1,3}.*`` / ``shift_magnitudes.{0,1,3}.*``.
Follows /root/reference/lib/reconstructor.py:71-79 and torchvision 0.26 ``resnet18`` (BasicBlock,
v1 stride placement).  BatchNorm runs in *train* mode (batch statistics, biased variance; running
statistics updated with momentum 0.1 and the unbiased variance), as lib/trainer.py:150 sets.
"""
import math
import torch
import torch.nn.functional as F

RESNET_STAGES = ((64, 1), (128, 2), (256, 2), (512, 2))    # (channels, stride of first block)


def _bn(sd, p, x, train=True, running=None, eps=1e-5, momentum=0.1):
    rm, rv = sd[p + '.running_mean'], sd[p + '.running_var']
    if train and running is not None:
        rm, rv = rm.clone(), rv.clone()
        running[p + '.running_mean'], running[p + '.running_var'] = rm, rv
    elif train:
        rm = rv = None
    return F.batch_norm(x, rm, rv, sd[p + '.weight'], sd[p + '.bias'], train, momentum, eps)


def resnet_features(sd, x, train=True, running=None, prefix='features_extractor'):
    """resnet18 up to (and including) the global average pool -> [B, 512]."""
    p = prefix
    x = F.conv2d(x, sd[p + '.conv1.weight'], None, 2, 3)
    x = F.relu(_bn(sd, p + '.bn1', x, train, running))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, (_, stride) in enumerate(RESNET_STAGES, start=1):
        for bi in range(2):
            q = '%s.layer%d.%d' % (p, li, bi)
            s = stride if bi == 0 else 1
            idt = x
            h = F.conv2d(x, sd[q + '.conv1.weight'], None, s, 1)
            h = F.relu(_bn(sd, q + '.bn1', h, train, running))
            h = F.conv2d(h, sd[q + '.conv2.weight'], None, 1, 1)
            h = _bn(sd, q + '.bn2', h, train, running)
            if (q + '.downsample.0.weight') in sd:
                idt = F.conv2d(x, sd[q + '.downsample.0.weight'], None, s, 0)
                idt = _bn(sd, q + '.downsample.1', idt, train, running)
            x = F.relu(h + idt)
    return x.mean(dim=[2, 3])


def lenet_features(sd, x, train=True, running=None, prefix='feature_extractor'):
    """lib/reconstructor.py:21-33,74: 3x (conv5x5, BN, ReLU[, maxpool2]) then spatial mean."""
    p = prefix
    x = F.conv2d(x, sd[p + '.0.weight'], sd[p + '.0.bias'])
    x = F.max_pool2d(F.relu(_bn(sd, p + '.1', x, train, running)), 2, 2)
    x = F.conv2d(x, sd[p + '.4.weight'], sd[p + '.4.bias'])
    x = F.max_pool2d(F.relu(_bn(sd, p + '.5', x, train, running)), 2, 2)
    x = F.conv2d(x, sd[p + '.8.weight'], sd[p + '.8.bias'])
    x = F.relu(_bn(sd, p + '.9', x, train, running))
    return x.mean(dim=[-1, -2]).view(x.shape[0], -1)


def _lenet_head(sd, p, f, train, running):
    h = F.linear(f, sd[p + '.0.weight'], sd[p + '.0.bias'])
    h = F.relu(_bn(sd, p + '.1', h, train, running))
    return F.linear(h, sd[p + '.3.weight'], sd[p + '.3.bias'])


def forward(sd, x1, x2, reconstructor_type='ResNet', train=True, running=None):
    """Reconstructor.forward: (path-index logits [B, K], shift magnitudes [B])."""
    x = torch.cat([x1, x2], dim=1)
    if reconstructor_type == 'ResNet':
        f = resnet_features(sd, x, train, running)
        logits = F.linear(f, sd['path_indices.weight'], sd['path_indices.bias'])
        mag = F.linear(f, sd['shift_magnitudes.weight'], sd['shift_magnitudes.bias'])
    elif reconstructor_type == 'LeNet':
        f = lenet_features(sd, x, train, running)
        logits = _lenet_head(sd, 'path_indices', f, train, running)
        mag = _lenet_head(sd, 'shift_magnitudes', f, train, running)
    else:
        raise ValueError(reconstructor_type)
    return logits, mag.squeeze()


def trainable_keys(sd):
    """Parameters that receive gradients (everything but BN buffers and the dead ``fc``)."""
    skip = ('running_mean', 'running_var', 'num_batches_tracked')
    return [k for k in sd if not k.endswith(skip) and '.fc.' not in k]


# ---------------------------------------------------------------------------------------------
# random init (torch defaults: kaiming-uniform(a=sqrt 5) for conv/linear, kaiming-normal fan_out for
# resnet convs, BN weight 1 / bias 0)
def _default_uniform(shape, g):
    fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(*shape, generator=g) * 2 - 1) * bound, bound


def _bn_init(sd, p, c):
    sd[p + '.weight'] = torch.ones(c)
    sd[p + '.bias'] = torch.zeros(c)
    sd[p + '.running_mean'] = torch.zeros(c)
    sd[p + '.running_var'] = torch.ones(c)
    sd[p + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def _linear_init(sd, p, out_f, in_f, g):
    w, bound = _default_uniform((out_f, in_f), g)
    sd[p + '.weight'] = w
    sd[p + '.bias'] = (torch.rand(out_f, generator=g) * 2 - 1) * bound


def init_state(reconstructor_type, dim, channels=3, generator=None):
    g = generator
    sd = {}
    if reconstructor_type == 'ResNet':
        p = 'features_extractor'

        def kconv(name, co, ci, k):
            sd[name] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(2.0 / (co * k * k))

        kconv(p + '.conv1.weight', 64, 6, 7)
        _bn_init(sd, p + '.bn1', 64)
        cin = 64
        for li, (c, stride) in enumerate(RESNET_STAGES, start=1):
            for bi in range(2):
                q = '%s.layer%d.%d' % (p, li, bi)
                kconv(q + '.conv1.weight', c, cin if bi == 0 else c, 3)
                _bn_init(sd, q + '.bn1', c)
                kconv(q + '.conv2.weight', c, c, 3)
                _bn_init(sd, q + '.bn2', c)
                if bi == 0 and (stride != 1 or cin != c):
                    kconv(q + '.downsample.0.weight', c, cin, 1)
                    _bn_init(sd, q + '.downsample.1', c)
            cin = c
        _linear_init(sd, p + '.fc', 1000, 512, g)
        _linear_init(sd, 'path_indices', dim, 512, g)
        _linear_init(sd, 'shift_magnitudes', 1, 512, g)
    elif reconstructor_type == 'LeNet':
        p = 'feature_extractor'
        widths = [(channels * 2, 6), (6, 16), (16, 120)]
        for idx, (ci, co) in zip((0, 4, 8), widths):
            w, bound = _default_uniform((co, ci, 5, 5), g)
            sd['%s.%d.weight' % (p, idx)] = w
            sd['%s.%d.bias' % (p, idx)] = (torch.rand(co, generator=g) * 2 - 1) * bound
            _bn_init(sd, '%s.%d' % (p, idx + 1), co)
        for head, out_f in (('path_indices', dim), ('shift_magnitudes', 1)):
            _linear_init(sd, head + '.0', 84, 120, g)
            _bn_init(sd, head + '.1', 84)
            _linear_init(sd, head + '.3', out_f, 84, g)
    else:
        raise ValueError(reconstructor_type)
    return sd
