"""CPU oracle: the attribute predictors of the attribute-space traversal, restated functionally over state dicts with the
reference's key names (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

  * FairFace: ``torchvision.models.resnet34`` with ``fc = Linear(512, 18)`` in eval mode (traverse_attribute_space.py:178-183)
    followed by the race / gender / age score arithmetic of :412-433;
  * Hopenet: ResNet-50 trunk + three 66-bin heads (lib/evaluation/hopenet/hopenet.py:5-66) followed by the soft-argmax pose of
    traverse_attribute_space.py:448-456.

  * the CelebA attribute predictor, the S3FD face detector, the ArcFace identity comparator (IR-SE-50) and the action-unit
    hourglass detector (sections below, each citing its reference file).

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



# ---- ArcFace identity comparator (lib/evaluation/archface/arcface.py:9-24,119-164) -------------------------------------------
ARCFACE_STAGES = ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3))          # get_blocks(50), arcface.py:103-108


def _bn_state(sd, name, c, generator, gain=1.0):
    sd[name + '.weight'] = gain * (0.5 + torch.rand(c, generator=generator))
    sd[name + '.bias'] = 0.1 * torch.randn(c, generator=generator)
    sd[name + '.running_mean'] = 0.1 * torch.randn(c, generator=generator)
    sd[name + '.running_var'] = 0.5 + torch.rand(c, generator=generator)
    sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def arcface_units():
    """(in_channel, depth, stride) of the 24 body units."""
    return [(ci if i == 0 else d, d, 2 if i == 0 else 1) for ci, d, n in ARCFACE_STAGES for i in range(n)]


def init_arcface_state(generator):
    """Seeded SE_IR(50, mode='ir_se') state with the reference's key names (BatchNorm statistics randomised)."""
    sd = {}

    def conv(name, co, ci, k, gain=1.0):
        sd[name + '.weight'] = torch.randn(co, ci, k, k, generator=generator) * math.sqrt(2.0 / (ci * k * k)) * gain

    conv('input_layer.0', 64, 3, 3)
    _bn_state(sd, 'input_layer.1', 64, generator)
    sd['input_layer.2.weight'] = 0.1 + 0.3 * torch.rand(64, generator=generator)
    for i, (ci, d, s) in enumerate(arcface_units()):
        p = 'body.%d' % i
        if ci != d:
            conv(p + '.shortcut_layer.0', d, ci, 1)
            _bn_state(sd, p + '.shortcut_layer.1', d, generator)
        _bn_state(sd, p + '.res_layer.0', ci, generator)
        conv(p + '.res_layer.1', d, ci, 3)
        sd[p + '.res_layer.2.weight'] = 0.1 + 0.3 * torch.rand(d, generator=generator)
        conv(p + '.res_layer.3', d, d, 3)
        _bn_state(sd, p + '.res_layer.4', d, generator, gain=0.5)
        conv(p + '.res_layer.5.fc1', d // 16, d, 1)
        conv(p + '.res_layer.5.fc2', d, d // 16, 1)
    _bn_state(sd, 'output_layer.0', 512, generator)
    sd['output_layer.3.weight'] = torch.randn(512, 512 * 7 * 7, generator=generator) / math.sqrt(512 * 7 * 7)
    sd['output_layer.3.bias'] = 0.05 * torch.randn(512, generator=generator)
    _bn_state(sd, 'output_layer.4', 512, generator)
    return sd


def arcface_backbone(sd, x):
    """SE_IR.forward (arcface.py:159-163) in eval mode: x [N, 3, 112, 112] -> unit-norm embeddings [N, 512]."""
    x = F.prelu(_bn(sd, 'input_layer.1', F.conv2d(x, sd['input_layer.0.weight'], None, 1, 1)), sd['input_layer.2.weight'])
    for i, (ci, d, s) in enumerate(arcface_units()):
        p = 'body.%d' % i
        if ci == d:
            sc = x[:, :, ::s, ::s]                                            # MaxPool2d(1, stride), :64-65
        else:
            sc = _bn(sd, p + '.shortcut_layer.1', F.conv2d(x, sd[p + '.shortcut_layer.0.weight'], None, s))
        r = F.conv2d(_bn(sd, p + '.res_layer.0', x), sd[p + '.res_layer.1.weight'], None, 1, 1)
        r = F.prelu(r, sd[p + '.res_layer.2.weight'])
        r = _bn(sd, p + '.res_layer.4', F.conv2d(r, sd[p + '.res_layer.3.weight'], None, s, 1))
        g = F.conv2d(F.relu(F.conv2d(r.mean(dim=(2, 3), keepdim=True), sd[p + '.res_layer.5.fc1.weight'])),
                     sd[p + '.res_layer.5.fc2.weight'])                       # SEModule, :51-58
        x = r * torch.sigmoid(g) + sc
    x = _bn(sd, 'output_layer.0', x).flatten(1)                               # Dropout is inactive in eval mode
    x = F.linear(x, sd['output_layer.3.weight'], sd['output_layer.3.bias'])
    x = F.batch_norm(x, sd['output_layer.4.running_mean'], sd['output_layer.4.running_var'], sd['output_layer.4.weight'],
                     sd['output_layer.4.bias'], False, 0.0, 1e-5)
    return x / torch.norm(x, 2, 1, True)


def arcface_extract_feats(sd, x):
    """IDComparator.extract_feats (arcface.py:16-19): fixed face region of a 256 x 256 frame, pooled to 112 x 112."""
    return arcface_backbone(sd, F.adaptive_avg_pool2d(x[:, :, 35:223, 32:220], (112, 112)))


def id_similarity(sd, x, x_prime):
    """IDComparator.forward (arcface.py:21-22)."""
    return F.cosine_similarity(arcface_extract_feats(sd, x), arcface_extract_feats(sd, x_prime), dim=1, eps=1e-6).mean()


# ---- Action-unit detector (lib/evaluation/au_detector/hourglass.py:17-243, AU_detector.py:29-46) ----------------------------
def _convblock_state(sd, name, ci, co, generator, lightweight=False):
    k = 1 if lightweight else 3
    for j, (a, b) in enumerate(((ci, co // 2), (co // 2, co // 4), (co // 4, co // 4)), start=1):
        sd['%s.conv%d.weight' % (name, j)] = torch.randn(b, a, k, k, generator=generator) * math.sqrt(2.0 / (a * k * k))
        _bn_state(sd, '%s.bn%d' % (name, j), b, generator)
    if ci != co:
        sd[name + '.downsample.0.weight'] = torch.randn(co, ci, 1, 1, generator=generator) * math.sqrt(2.0 / ci)
        _bn_state(sd, name + '.downsample.1', co, generator)


def _hourglass_blocks(depth=4):
    """Module names of HourGlass(1, depth, .) in registration order, with the lightweight flag of b1 (hourglass.py:78-89)."""
    names = []

    def gen_(level):
        names.append(('b1_%d' % level, True))
        names.append(('b2_%d' % level, False))
        if level > 1:
            gen_(level - 1)
        else:
            names.append(('b2_plus_%d' % level, False))
        names.append(('b3_%d' % level, False))

    gen_(depth)
    return names


def init_au_state(generator, n_points=12):
    """Seeded FANAU(num_modules=1, n_points=12) state with the reference's key names."""
    sd = {}

    def conv_b(name, co, ci, k):
        sd[name + '.weight'] = torch.randn(co, ci, k, k, generator=generator) * math.sqrt(2.0 / (ci * k * k))
        sd[name + '.bias'] = 0.05 * torch.randn(co, generator=generator)

    conv_b('fan.conv1', 64, 3, 7)
    _bn_state(sd, 'fan.bn1', 64, generator)
    _convblock_state(sd, 'fan.conv2', 64, 64, generator)
    _convblock_state(sd, 'fan.conv3', 64, 128, generator)
    _convblock_state(sd, 'fan.conv4', 128, 128, generator)
    for name, _ in _hourglass_blocks():
        _convblock_state(sd, 'fan.m0.' + name, 128, 128, generator)          # the landmark hourglass has no lightweight blocks
    _convblock_state(sd, 'fan.top_m_0', 128, 128, generator)
    conv_b('fan.conv_last0', 128, 128, 1)
    _bn_state(sd, 'fan.bn_end0', 128, generator)
    conv_b('fan.l0', 68, 128, 1)
    conv_b('conv1.0', 128, 68, 1)
    _bn_state(sd, 'conv1.1', 128, generator)
    conv_b('conv2.0', 128, 128, 1)
    _bn_state(sd, 'conv2.1', 128, generator)
    for name, light in _hourglass_blocks():
        _convblock_state(sd, 'net.' + name, 128, 128, generator, lightweight=light)
    conv_b('conv_last.0', 128, 128, 1)
    _bn_state(sd, 'conv_last.1', 128, generator)
    conv_b('l', n_points, 128, 1)
    return sd


def _convblock(sd, name, x):
    """ConvBlock.forward (hourglass.py:46-68); kernel size and padding follow the stored weights (lightweight = 1 x 1)."""
    pad = (sd[name + '.conv1.weight'].shape[-1] - 1) // 2
    o1 = F.relu6(_bn(sd, name + '.bn1', F.conv2d(x, sd[name + '.conv1.weight'], None, 1, pad)))
    o2 = F.relu6(_bn(sd, name + '.bn2', F.conv2d(o1, sd[name + '.conv2.weight'], None, 1, pad)))
    o3 = F.relu6(_bn(sd, name + '.bn3', F.conv2d(o2, sd[name + '.conv3.weight'], None, 1, pad)))
    res = x
    if name + '.downsample.0.weight' in sd:
        res = F.relu6(_bn(sd, name + '.downsample.1', F.conv2d(x, sd[name + '.downsample.0.weight'])))
    return torch.cat((o1, o2, o3), 1) + res


def _hourglass(sd, name, level, x):
    """HourGlass._forward (hourglass.py:91-113)."""
    up1 = _convblock(sd, '%s.b1_%d' % (name, level), x)
    low = _convblock(sd, '%s.b2_%d' % (name, level), F.max_pool2d(x, 2, 2))
    low = _hourglass(sd, name, level - 1, low) if level > 1 else _convblock(sd, '%s.b2_plus_%d' % (name, level), low)
    low = _convblock(sd, '%s.b3_%d' % (name, level), low)
    return up1 + F.interpolate(low, scale_factor=2, mode='nearest')


def au_heatmaps(sd, x):
    """FANAU.forward (hourglass.py:224-243) over QFAN.forward (:154-185) in eval mode: x [N, 3, 256, 256] -> [N, 12, 64, 64]."""
    h = F.relu(_bn(sd, 'fan.bn1', F.conv2d(x, sd['fan.conv1.weight'], sd['fan.conv1.bias'], 2, 3)))
    h = F.max_pool2d(_convblock(sd, 'fan.conv2', h), 2, 2)
    feat = _convblock(sd, 'fan.conv4', _convblock(sd, 'fan.conv3', h))
    ll = _convblock(sd, 'fan.top_m_0', _hourglass(sd, 'fan.m0', 4, feat))
    ll = F.relu(_bn(sd, 'fan.bn_end0', F.conv2d(ll, sd['fan.conv_last0.weight'], sd['fan.conv_last0.bias'])))
    lmk = F.conv2d(ll, sd['fan.l0.weight'], sd['fan.l0.bias'])
    a = F.relu6(_bn(sd, 'conv1.1', F.conv2d(lmk, sd['conv1.0.weight'], sd['conv1.0.bias'])))
    b = F.relu6(_bn(sd, 'conv2.1', F.conv2d(feat, sd['conv2.0.weight'], sd['conv2.0.bias'])))
    h = _hourglass(sd, 'net', 4, a + b)
    h = F.relu6(_bn(sd, 'conv_last.1', F.conv2d(h, sd['conv_last.0.weight'], sd['conv_last.0.bias'])))
    return F.conv2d(h, sd['l.weight'], sd['l.bias'])


def detect_au(sd, img):
    """AUdetector.detect_AU (AU_detector.py:35-46): min-max normalised batch -> heat-map maxima [N, 12]."""
    x = (img - img.min()) / (img.max() - img.min())
    return F.max_pool2d(au_heatmaps(sd, x), (64, 64)).squeeze(2).squeeze(2)
