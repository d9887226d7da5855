"""Frozen torchvision-style ResNets for the attribute-space traversal.

Two of the reference's attribute predictors are plain ImageNet ResNets run in eval mode on 224 x 224 face crops
(traverse_attribute_space.py:178-196): FairFace = ``torchvision.models.resnet34`` with an 18-way ``fc`` (race 7, gender 2,
age 9) and Hopenet = a ResNet-50 trunk with three 66-bin heads (lib/evaluation/hopenet/hopenet.py:5-66).  Here they are
inference-only kernel chains on libwgs_b200:

  * eval-mode BatchNorm is folded into the conv that feeds it once (``plan()``): w' = w * gamma / sqrt(var + eps),
    b' = beta - mean * gamma / sqrt(var + eps);
  * every conv is one tensor-core launch whose epilogue adds the bias, adds the residual (``accumulate`` into the buffer
    that already holds the identity / down-sample branch), applies ReLU and writes the next conv's split32 operand
    (``out_split``) - no separate BatchNorm, add, ReLU or pack kernels;
  * the 7x7 / 2 stem on 3 channels runs as a stride-1 4x4-tap conv over the 2x2 space-to-depth input, like the
    Reconstructor's.

The parameter tree carries torchvision's names (``conv1.weight``, ``layer2.0.downsample.1.running_var``, ``fc.weight`` /
``fc_yaw.weight`` ...), so the published checkpoints load with ``load_state_dict``.  CUDA only, no CPU fallback.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import _lib
from . import conv as C


class _Holder(nn.Module):
    """Parameter container (compute lives in FrozenResNet.forward)."""


def _conv(co, ci, k):
    m = _Holder()
    w = torch.empty(co, ci, k, k)
    nn.init.kaiming_normal_(w, mode='fan_out', nonlinearity='relu')
    m.weight = nn.Parameter(w, requires_grad=False)
    return m


def _bn(c):
    m = nn.BatchNorm2d(c)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


class FrozenResNet(nn.Module):
    """block: 'basic' (expansion 1) or 'bottleneck' (expansion 4); layers: blocks per stage; heads: {name: out_features}
    linear heads on the pooled features (``{'fc': 1000}`` = torchvision; ``fc_yaw / fc_pitch / fc_roll`` = Hopenet)."""

    def __init__(self, block, layers, heads):
        super().__init__()
        assert block in ('basic', 'bottleneck')
        self.block, self.expansion = block, 1 if block == 'basic' else 4
        self.conv1, self.bn1 = _conv(64, 3, 7), _bn(64)
        inplanes = 64
        self.stage_names = []
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            blocks = []
            for bi in range(n):
                stride = 2 if (bi == 0 and li > 1) else 1
                b = _Holder()
                b.stride = stride
                if block == 'basic':
                    b.conv1, b.bn1 = _conv(planes, inplanes, 3), _bn(planes)
                    b.conv2, b.bn2 = _conv(planes, planes, 3), _bn(planes)
                else:       # torchvision >= 0.3 / Hopenet's Bottleneck: the stride sits on the 3x3 conv
                    b.conv1, b.bn1 = _conv(planes, inplanes, 1), _bn(planes)
                    b.conv2, b.bn2 = _conv(planes, planes, 3), _bn(planes)
                    b.conv3, b.bn3 = _conv(planes * 4, planes, 1), _bn(planes * 4)
                if stride != 1 or inplanes != planes * self.expansion:
                    b.downsample = nn.Sequential(_conv(planes * self.expansion, inplanes, 1), _bn(planes * self.expansion))
                inplanes = planes * self.expansion
                blocks.append(b)
            setattr(self, 'layer%d' % li, nn.Sequential(*blocks))
        self.head_names = list(heads)
        for name, out in heads.items():
            setattr(self, name, nn.Linear(inplanes, out))
        for p in self.parameters():
            p.requires_grad_(False)
        self._plan = None
        self.eval()

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _fold(conv, bn):
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        return (conv.weight * s.view(-1, 1, 1, 1)).float().contiguous(), (bn.bias - bn.running_mean * s).float().contiguous()

    def plan(self):
        if self._plan is not None:
            return self._plan
        if self.conv1.weight.device.type != 'cuda':
            raise RuntimeError('FrozenResNet runs on CUDA only (no CPU fallback); call .cuda() first')
        P = {}
        with torch.no_grad():
            w, b = self._fold(self.conv1, self.bn1)
            P['stem'] = (C.pack_weights_group([(w, C.PACK_S2D, 3)])[0], b)
            P['blocks'] = []
            for li in range(1, 5):
                for blk in getattr(self, 'layer%d' % li):
                    e = dict(stride=blk.stride, convs=[])
                    names = (('conv1', 'bn1'), ('conv2', 'bn2')) + ((('conv3', 'bn3'),) if self.block == 'bottleneck' else ())
                    for cn, bn in names:
                        w, b = self._fold(getattr(blk, cn), getattr(blk, bn))
                        e['convs'].append((C.pack_weights(w), b, w.shape))
                    if hasattr(blk, 'downsample'):
                        w, b = self._fold(blk.downsample[0], blk.downsample[1])
                        e['down'] = (C.pack_weights(w), b, w.shape)
                    P['blocks'].append(e)
        self._plan = P
        return P

    def features(self, x):
        """x [N, 3, H, W] (H, W even) -> pooled features [N, 512 * expansion]."""
        if not x.is_cuda:
            raise RuntimeError('FrozenResNet runs on CUDA tensors only (no CPU fallback); got %s' % x.device)
        P = self.plan()
        n, _, h, w = x.shape
        dev = x.device
        x_nhwc = x.float().permute(0, 2, 3, 1).contiguous()
        w_stem, b_stem = P['stem']
        oh, ow = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
        taps, _ = C.s2d_taps(7, 3)
        y0 = torch.empty(n, oh, ow, 64, device=dev, dtype=torch.float32)
        C.conv_taps(C.s2d_pack_split32(x_nhwc), w_stem, taps, y0, grid=(oh, ow), cin=7 * 7 * 3, beta=b_stem, act=1,
                    algo_macs_per_pixel=7 * 7 * 3 * 64)
        ph, pw = (oh + 1) // 2, (ow + 1) // 2
        cur = torch.empty(n, ph, pw, 64, device=dev, dtype=torch.float32)
        idx = torch.empty(n, ph, pw, 64, device=dev, dtype=torch.uint8)
        cur_s = torch.empty(n, ph, pw, 2, 64, device=dev, dtype=torch.bfloat16)
        _lib.call('wgs_maxpool3s2_fwd', _lib.ptr(y0), n, oh, ow, 64, _lib.ptr(cur), _lib.ptr(idx), _lib.ptr(cur_s), _lib.stream())

        def conv(xs, packed, bias, shape, stride, out=None, accumulate=False, want_f32=False, want_split=True):
            co, ci, k, _ = shape
            pad = (k - 1) // 2
            oh_, ow_ = (xs.shape[1] + 2 * pad - k) // stride + 1, (xs.shape[2] + 2 * pad - k) // stride + 1
            split = torch.empty(n, oh_, ow_, C.chunks_of(co), 64, device=dev, dtype=torch.bfloat16) if want_split else None
            if out is None and want_f32:
                out = torch.empty(n, oh_, ow_, co, device=dev, dtype=torch.float32)
            C.conv2d(xs, packed, k, k, stride=stride, padding=pad, out=out, no_f32=out is None, cin=ci, beta=bias, act=1,
                     accumulate=accumulate, out_split=split)
            return out, split

        for e in P['blocks']:
            if 'down' in e:         # the shortcut branch first: its (bias-only, no ReLU) output is what the last conv adds to
                pk, b, shape = e['down']
                co = shape[0]
                oh_, ow_ = (cur_s.shape[1] - 1) // e['stride'] + 1, (cur_s.shape[2] - 1) // e['stride'] + 1
                res = torch.empty(n, oh_, ow_, co, device=dev, dtype=torch.float32)
                C.conv2d(cur_s, pk, 1, 1, stride=e['stride'], padding=0, out=res, cin=shape[1], beta=b, act=0)
            else:
                res = cur           # accumulated in place: the block input is not needed again
            xs = cur_s
            strides = (e['stride'], 1) if self.block == 'basic' else (1, e['stride'], 1)
            for (pk, b, shape), st in list(zip(e['convs'], strides))[:-1]:
                _, xs = conv(xs, pk, b, shape, st)
            pk, b, shape = e['convs'][-1]
            cur, cur_s = conv(xs, pk, b, shape, strides[-1], out=res, accumulate=True)
        return cur.mean(dim=(1, 2))

    def forward(self, x):
        f = self.features(x)
        outs = tuple(F.linear(f, getattr(self, h).weight, getattr(self, h).bias) for h in self.head_names)
        return outs[0] if len(outs) == 1 else outs


def fairface_resnet34():
    """traverse_attribute_space.py:178-183: resnet34 with an 18-way fc (race [0:7], gender [7:9], age [9:18])."""
    return FrozenResNet('basic', (3, 4, 6, 3), {'fc': 18})


def hopenet_resnet50(num_bins=66):
    """traverse_attribute_space.py:186 / lib/evaluation/hopenet/hopenet.py:5-66 (the vestigial fc_finetune is kept for the
    state dict and never evaluated, as in the reference)."""
    m = FrozenResNet('bottleneck', (3, 4, 6, 3), {'fc_yaw': num_bins, 'fc_pitch': num_bins, 'fc_roll': num_bins})
    m.fc_finetune = nn.Linear(512 * 4 + 3, 3)
    for p in m.fc_finetune.parameters():
        p.requires_grad_(False)
    return m


# attributes_5.json of the reference (lib/evaluation/celeba_attributes/attributes_5.json): CelebA attribute id -> (name, classes)
CELEBA_5 = (('6', 'Bangs', 6), ('16', 'Eyeglasses', 6), ('25', 'No_Beard', 6), ('32', 'Smiling', 6), ('40', 'Young', 6))


class _FcBlock(nn.Module):
    """celeba_attr_predictor.py:86-103 in eval mode: Linear -> BatchNorm1d -> ReLU (the drop-out is inactive)."""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.fc = nn.Linear(inplanes, planes)
        self.bn = nn.BatchNorm1d(planes)

    def forward(self, x):
        return F.relu(self.bn(self.fc(x)))


class CelebAPredictor(FrozenResNet):
    """lib/evaluation/celeba_attributes/celeba_attr_predictor.py:106-182: ResNet-50 trunk (the kernel chain above), a
    2048 -> 512 ``stem`` block and one small classifier per attribute (``classifier06Bangs`` ...); forward returns
    {attribute name: logits}.  Same state-dict keys as the reference's ``ResNet(Bottleneck, [3, 4, 6, 3], attr_file)``."""

    def __init__(self, attr_info=CELEBA_5):
        super().__init__('bottleneck', (3, 4, 6, 3), {})
        self.attr_info = tuple(attr_info)
        self.stem = _FcBlock(512 * 4, 512)
        for key, name, num in self.attr_info:
            setattr(self, 'classifier' + str(key).zfill(2) + name, nn.Sequential(_FcBlock(512, 256), nn.Linear(256, num)))
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def forward(self, x):
        f = self.stem(self.features(x))
        return {name: getattr(self, 'classifier' + str(key).zfill(2) + name)(f) for key, name, _ in self.attr_info}


def celeba_attr_resnet50(attr_info=CELEBA_5):
    """traverse_attribute_space.py:206-207."""
    return CelebAPredictor(attr_info)

