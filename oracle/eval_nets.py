"""CPU oracle: the two ResNet-based attribute predictors of the attribute-space traversal, restated functionally over
torchvision-named state dicts (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

  * FairFace: ``torchvision.models.resnet34`` with ``fc = Linear(512, 18)`` in eval mode (traverse_attribute_space.py:178-183)
    followed by the race / gender / age score arithmetic of :412-433;
  * Hopenet: ResNet-50 trunk + three 66-bin heads (lib/evaluation/hopenet/hopenet.py:5-66) followed by the soft-argmax pose of
    traverse_attribute_space.py:448-456.

Third-party arithmetic: torchvision's BasicBlock / Bottleneck (no version pinned by the reference's requirements.txt; 0.26.0
installed here).  Pinned by oracle/gen_golden.py::pin_eval_nets against torchvision.models.resnet34 and the reference's own
Hopenet class on seeded weights with randomised BatchNorm statistics; fixture tests/golden/eval_nets.pt.
"""
import math

import torch
import torch.nn.functional as F

LAYERS = (3, 4, 6, 3)


def init_state(block, heads, generator, layers=LAYERS):
    """Seeded random state with torchvision's key names; BatchNorm statistics are randomised so that eval-mode BatchNorm is
    not the identity.  block: 'basic' | 'bottleneck'; heads: {name: out_features}."""
    exp = 1 if block == 'basic' else 4
    sd = {}

    def conv(name, co, ci, k):
        sd[name + '.weight'] = torch.randn(co, ci, k, k, generator=generator) * math.sqrt(2.0 / (k * k * co))

    def bn(name, c):
        sd[name + '.weight'] = 0.5 + torch.rand(c, generator=generator)
        sd[name + '.bias'] = 0.1 * torch.randn(c, generator=generator)
        sd[name + '.running_mean'] = 0.1 * torch.randn(c, generator=generator)
        sd[name + '.running_var'] = 0.5 + torch.rand(c, generator=generator)
        sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    conv('conv1', 64, 3, 7)
    bn('bn1', 64)
    inplanes = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for bi in range(n):
            p = 'layer%d.%d' % (li, bi)
            stride = 2 if (bi == 0 and li > 1) else 1
            if block == 'basic':
                conv(p + '.conv1', planes, inplanes, 3); bn(p + '.bn1', planes)
                conv(p + '.conv2', planes, planes, 3); bn(p + '.bn2', planes)
            else:
                conv(p + '.conv1', planes, inplanes, 1); bn(p + '.bn1', planes)
                conv(p + '.conv2', planes, planes, 3); bn(p + '.bn2', planes)
                conv(p + '.conv3', planes * 4, planes, 1); bn(p + '.bn3', planes * 4)
            if stride != 1 or inplanes != planes * exp:
                conv(p + '.downsample.0', planes * exp, inplanes, 1); bn(p + '.downsample.1', planes * exp)
            inplanes = planes * exp
    for name, out in heads.items():
        bound = 1.0 / math.sqrt(inplanes)
        sd[name + '.weight'] = (torch.rand(out, inplanes, generator=generator) * 2 - 1) * bound
        sd[name + '.bias'] = (torch.rand(out, generator=generator) * 2 - 1) * bound
    return sd


def _bn(sd, name, x, eps=1e-5):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'], sd[name + '.bias'],
                        False, 0.0, eps)


def resnet_forward(sd, x, block, heads, layers=LAYERS):
    """Eval-mode forward (torchvision/models/resnet.py ResNet._forward_impl; Hopenet.forward, hopenet.py:51-66).
    Returns the tuple of head outputs in `heads` order."""
    f = resnet_features(sd, x, block, layers)
    return tuple(F.linear(f, sd[h + '.weight'], sd[h + '.bias']) for h in heads)


def resnet_features(sd, x, block, layers=LAYERS):
    """Pooled trunk features [N, 512 * expansion]."""
    x = F.max_pool2d(F.relu(_bn(sd, 'bn1', F.conv2d(x, sd['conv1.weight'], None, 2, 3))), 3, 2, 1)
    for li, n in enumerate(layers, start=1):
        for bi in range(n):
            p = 'layer%d.%d' % (li, bi)
            stride = 2 if (bi == 0 and li > 1) else 1
            idt = x
            if block == 'basic':
                h = F.relu(_bn(sd, p + '.bn1', F.conv2d(x, sd[p + '.conv1.weight'], None, stride, 1)))
                h = _bn(sd, p + '.bn2', F.conv2d(h, sd[p + '.conv2.weight'], None, 1, 1))
            else:
                h = F.relu(_bn(sd, p + '.bn1', F.conv2d(x, sd[p + '.conv1.weight'])))
                h = F.relu(_bn(sd, p + '.bn2', F.conv2d(h, sd[p + '.conv2.weight'], None, stride, 1)))
                h = _bn(sd, p + '.bn3', F.conv2d(h, sd[p + '.conv3.weight']))
            if p + '.downsample.0.weight' in sd:
                idt = _bn(sd, p + '.downsample.1', F.conv2d(x, sd[p + '.downsample.0.weight'], None, stride, 0))
            x = F.relu(h + idt)
    return x.mean(dim=(2, 3))       # AvgPool2d(7) on the 7 x 7 map of a 224 x 224 crop == AdaptiveAvgPool2d(1)


CELEBA_5 = (('6', 'Bangs', 6), ('16', 'Eyeglasses', 6), ('25', 'No_Beard', 6), ('32', 'Smiling', 6), ('40', 'Young', 6))


def init_celeba_state(generator, attr_info=CELEBA_5):
    """State of the CelebA attribute predictor (celeba_attr_predictor.py:106-135): ResNet-50 trunk + stem + classifiers."""
    sd = init_state('bottleneck', {}, generator)

    def fc_block(name, i, o):
        sd[name + '.fc.weight'] = (torch.rand(o, i, generator=generator) * 2 - 1) / math.sqrt(i)
        sd[name + '.fc.bias'] = (torch.rand(o, generator=generator) * 2 - 1) / math.sqrt(i)
        sd[name + '.bn.weight'] = 0.5 + torch.rand(o, generator=generator)
        sd[name + '.bn.bias'] = 0.1 * torch.randn(o, generator=generator)
        sd[name + '.bn.running_mean'] = 0.1 * torch.randn(o, generator=generator)
        sd[name + '.bn.running_var'] = 0.5 + torch.rand(o, generator=generator)
        sd[name + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    fc_block('stem', 2048, 512)
    for key, name, num in attr_info:
        c = 'classifier' + str(key).zfill(2) + name
        fc_block(c + '.0', 512, 256)
        sd[c + '.1.weight'] = (torch.rand(num, 256, generator=generator) * 2 - 1) / 16.0
        sd[c + '.1.bias'] = (torch.rand(num, generator=generator) * 2 - 1) / 16.0
    return sd


def celeba_forward(sd, x, attr_info=CELEBA_5):
    """celeba_attr_predictor.py:163-182 in eval mode -> {attribute name: logits}."""
    f = resnet_features(sd, x, 'bottleneck')

    def fc_block(name, t):
        t = F.linear(t, sd[name + '.fc.weight'], sd[name + '.fc.bias'])
        return F.relu(F.batch_norm(t, sd[name + '.bn.running_mean'], sd[name + '.bn.running_var'], sd[name + '.bn.weight'],
                                   sd[name + '.bn.bias'], False, 0.0, 1e-5))

    f = fc_block('stem', f)
    out = {}
    for key, name, _ in attr_info:
        c = 'classifier' + str(key).zfill(2) + name
        out[name] = F.linear(fc_block(c + '.0', f), sd[c + '.1.weight'], sd[c + '.1.bias'])
    return out


def fairface_scores(outputs):
    """traverse_attribute_space.py:412-433 -> (femaleness, age, race) score vectors."""
    o = outputs.double()
    gender = torch.softmax(o[:, 7:9], dim=1)[:, 1]
    age_s = torch.softmax(o[:, 9:18], dim=1)
    age = (age_s.argmax(dim=1) + age_s.max(dim=1).values) / 9.0
    race_s = torch.softmax(o[:, :7], dim=1)
    race = (race_s.argmax(dim=1) + race_s.max(dim=1).values) / 7.0
    return gender, age, race


def hopenet_pose(yaw, pitch, roll):
    """traverse_attribute_space.py:448-456: expectation over the 66 bins * 3 - 99, in degrees."""
    idx = torch.arange(66, dtype=yaw.dtype)
    return tuple(torch.sum(torch.softmax(t, dim=1) * idx, 1) * 3 - 99 for t in (yaw, pitch, roll))


# ---- S3FD face detector (lib/evaluation/sfd/net_s3fd.py:21-129) ---------------------------------------------------------------
S3FD_TRUNK = (('conv1_1', 3, 64, 3, 1, 1), ('conv1_2', 64, 64, 3, 1, 1), '|', ('conv2_1', 64, 128, 3, 1, 1), ('conv2_2', 128, 128, 3, 1, 1), '|',
              ('conv3_1', 128, 256, 3, 1, 1), ('conv3_2', 256, 256, 3, 1, 1), ('conv3_3', 256, 256, 3, 1, 1), '|',
              ('conv4_1', 256, 512, 3, 1, 1), ('conv4_2', 512, 512, 3, 1, 1), ('conv4_3', 512, 512, 3, 1, 1), '|',
              ('conv5_1', 512, 512, 3, 1, 1), ('conv5_2', 512, 512, 3, 1, 1), ('conv5_3', 512, 512, 3, 1, 1), '|',
              ('fc6', 512, 1024, 3, 1, 3), ('fc7', 1024, 1024, 1, 1, 0), ('conv6_1', 1024, 256, 1, 1, 0), ('conv6_2', 256, 512, 3, 2, 1),
              ('conv7_1', 512, 128, 1, 1, 0), ('conv7_2', 128, 256, 3, 2, 1))
S3FD_HEADS = (('conv3_3_norm', 'conv3_3', 256, 4, 10.0), ('conv4_3_norm', 'conv4_3', 512, 2, 8.0), ('conv5_3_norm', 'conv5_3', 512, 2, 5.0),
              ('fc7', 'fc7', 1024, 2, None), ('conv6_2', 'conv6_2', 512, 2, None), ('conv7_2', 'conv7_2', 256, 2, None))


def init_s3fd_state(generator):
    sd = {}

    def conv(name, ci, co, k, wscale=1.0):
        sd[name + '.weight'] = torch.randn(co, ci, k, k, generator=generator) * math.sqrt(2.0 / (ci * k * k)) * wscale
        sd[name + '.bias'] = 0.05 * torch.randn(co, generator=generator)

    for spec in S3FD_TRUNK:
        if spec != '|':
            conv(spec[0], spec[1], spec[2], spec[3])
    for head, _, c, ncls, scale in S3FD_HEADS:
        if scale is not None:
            sd[head + '.weight'] = scale * (0.8 + 0.4 * torch.rand(c, generator=generator))
        conv(head + '_mbox_conf', c, ncls, 3, wscale=0.6)       # spread-out face scores: some positions above 0.5, most below
        sd[head + '_mbox_conf.bias'][-1] -= 2.0                  # (the face class is the last channel)
        conv(head + '_mbox_loc', c, 4, 3, wscale=0.3)
    return sd


def s3fd_forward(sd, x):
    """net_s3fd.py:71-129 -> [cls1, reg1, ..., cls6, reg6]."""
    taps, h = {}, x
    for spec in S3FD_TRUNK:
        if spec == '|':
            h = F.max_pool2d(h, 2, 2)
            continue
        name, _, _, _, s, p = spec
        h = F.relu(F.conv2d(h, sd[name + '.weight'], sd[name + '.bias'], s, p))
        taps[name] = h
    outs = []
    for head, src, _, _, scale in S3FD_HEADS:
        f = taps[src]
        if scale is not None:                                     # L2Norm, :9-19
            f = f / (f.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10) * sd[head + '.weight'].view(1, -1, 1, 1)
        outs.append(F.conv2d(f, sd[head + '_mbox_conf.weight'], sd[head + '_mbox_conf.bias'], 1, 1))
        outs.append(F.conv2d(f, sd[head + '_mbox_loc.weight'], sd[head + '_mbox_loc.bias'], 1, 1))
    c = torch.chunk(outs[0], 4, 1)
    outs[0] = torch.cat([torch.max(torch.max(c[0], c[1]), c[2]), c[3]], dim=1)
    return outs


def sfd_detect_from_batch(olist):
    """sfd_detector.py:23-40 + detect.py:26-62 + bbox.py:48-66,94-111 restated per image: soft-max, positions with face score >
    0.05, anchor decode (stride 2^(i+2), anchor 4 x stride, variances 0.1 / 0.2), NMS 0.3, keep score > 0.5.
    Returns per-image [M_j, 5] arrays.  (The reference gathers the candidate positions over the whole batch for every image;
    the extra low-score entries are removed again by its NMS + 0.5 filter - pinned in gen_golden.pin_eval_nets.)"""
    import numpy as np
    res = []
    for j in range(olist[0].shape[0]):
        rows = []
        for i in range(len(olist) // 2):
            ocls = torch.softmax(olist[i * 2][j: j + 1], dim=1)[0, 1]
            oreg = olist[i * 2 + 1][j]
            stride = 2 ** (i + 2)
            hh, ww = torch.nonzero(ocls > 0.05, as_tuple=True)
            for h_, w_ in zip(hh.tolist(), ww.tolist()):
                axc, ayc, a = stride / 2 + w_ * stride, stride / 2 + h_ * stride, stride * 4.0
                loc = oreg[:, h_, w_]
                xc, yc = axc + float(loc[0]) * 0.1 * a, ayc + float(loc[1]) * 0.1 * a
                bw, bh = a * math.exp(float(loc[2]) * 0.2), a * math.exp(float(loc[3]) * 0.2)
                rows.append([xc - bw / 2, yc - bh / 2, xc - bw / 2 + bw, yc - bh / 2 + bh, float(ocls[h_, w_])])
        d = np.array(rows, dtype=np.float64).reshape(-1, 5)
        keep = []
        order = d[:, 4].argsort()[::-1]
        areas = (d[:, 2] - d[:, 0] + 1) * (d[:, 3] - d[:, 1] + 1)
        while order.size > 0:
            i0 = order[0]
            keep.append(i0)
            rest = order[1:]
            w_ = np.maximum(0.0, np.minimum(d[i0, 2], d[rest, 2]) - np.maximum(d[i0, 0], d[rest, 0]) + 1)
            h_ = np.maximum(0.0, np.minimum(d[i0, 3], d[rest, 3]) - np.maximum(d[i0, 1], d[rest, 1]) + 1)
            ovr = w_ * h_ / (areas[i0] + areas[rest] - w_ * h_)
            order = rest[ovr <= 0.3]
        kept = d[keep] if keep else d
        res.append(kept[kept[:, 4] > 0.5])
    return res

