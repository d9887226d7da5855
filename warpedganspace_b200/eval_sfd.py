"""S3FD face detector of the attribute-space traversal (reference: lib/evaluation/sfd/net_s3fd.py:21-129,
sfd_detector.py:5-43, detect.py:26-62, bbox.py:48-111) on libwgs_b200.

The network is a VGG-16 style stack of 3x3 convolutions with bias + ReLU, five 2x2 max-pools, three L2-normalised feature
taps and twelve small detection heads.  Every convolution is one tensor-core launch whose epilogue adds the bias, applies the
ReLU and - where the next layer is another convolution - writes that layer's split32 operand; max-pool, L2Norm and the
soft-max / box decoding of the post-processing are torch library calls on small tensors, batched over the frames of a path, and
the NMS is torchvision's CUDA op (on boxes shifted to the reference's pixel-count convention).  Parameter names are the
reference's (``conv1_1.weight`` ... ``conv7_2_mbox_loc.bias``), so ``s3fd-619a316812.pth`` loads as it is.  CUDA only.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import conv as C

# (name, cin, cout, kernel, stride, padding) in forward order; '|' marks a 2x2 max-pool, names in TAPS are feature taps
_TRUNK = (('conv1_1', 3, 64, 3, 1, 1), ('conv1_2', 64, 64, 3, 1, 1), '|',
          ('conv2_1', 64, 128, 3, 1, 1), ('conv2_2', 128, 128, 3, 1, 1), '|',
          ('conv3_1', 128, 256, 3, 1, 1), ('conv3_2', 256, 256, 3, 1, 1), ('conv3_3', 256, 256, 3, 1, 1), '|',
          ('conv4_1', 256, 512, 3, 1, 1), ('conv4_2', 512, 512, 3, 1, 1), ('conv4_3', 512, 512, 3, 1, 1), '|',
          ('conv5_1', 512, 512, 3, 1, 1), ('conv5_2', 512, 512, 3, 1, 1), ('conv5_3', 512, 512, 3, 1, 1), '|',
          ('fc6', 512, 1024, 3, 1, 3), ('fc7', 1024, 1024, 1, 1, 0),
          ('conv6_1', 1024, 256, 1, 1, 0), ('conv6_2', 256, 512, 3, 2, 1),
          ('conv7_1', 512, 128, 1, 1, 0), ('conv7_2', 128, 256, 3, 2, 1))
_TAPS = ('conv3_3', 'conv4_3', 'conv5_3', 'fc7', 'conv6_2', 'conv7_2')
_NORMS = {'conv3_3': ('conv3_3_norm', 256, 10.0), 'conv4_3': ('conv4_3_norm', 512, 8.0), 'conv5_3': ('conv5_3_norm', 512, 5.0)}
_HEADS = (('conv3_3_norm', 256, 4), ('conv4_3_norm', 512, 2), ('conv5_3_norm', 512, 2), ('fc7', 1024, 2), ('conv6_2', 512, 2),
          ('conv7_2', 256, 2))


class _L2Norm(nn.Module):
    def __init__(self, c, scale):
        super().__init__()
        self.weight = nn.Parameter(torch.full((c,), float(scale)), requires_grad=False)
        self.eps = 1e-10


class S3FD(nn.Module):
    """net_s3fd.s3fd: forward(x [N, 3, H, W] float, 0..255) -> [cls1, reg1, ..., cls6, reg6] (NCHW, like the reference)."""

    def __init__(self):
        super().__init__()
        for spec in _TRUNK:
            if spec != '|':
                name, ci, co, k, s, p = spec
                setattr(self, name, nn.Conv2d(ci, co, k, s, p))
        for name, c, scale in _NORMS.values():
            setattr(self, name, _L2Norm(c, scale))
        for src, c, ncls in _HEADS:
            setattr(self, src + '_mbox_conf', nn.Conv2d(c, ncls, 3, 1, 1))
            setattr(self, src + '_mbox_loc', nn.Conv2d(c, 4, 3, 1, 1))
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

    def plan(self):
        if self._plan is None:
            if self.conv1_1.weight.device.type != 'cuda':
                raise RuntimeError('S3FD runs on CUDA only (no CPU fallback); call .cuda() first')
            with torch.no_grad():
                self._plan = {n: (C.pack_weights(m.weight.float().contiguous()), m.bias.float().contiguous())
                              for n, m in self.named_children() if isinstance(m, nn.Conv2d)}
        return self._plan

    def _conv(self, name, xs, act, split, f32):
        m = getattr(self, name)
        w, b = self.plan()[name]
        k, s, p = m.kernel_size[0], m.stride[0], m.padding[0]
        n, h, wd = xs.shape[0], xs.shape[1], xs.shape[2]
        oh, ow = (h + 2 * p - k) // s + 1, (wd + 2 * p - k) // s + 1
        out_split = torch.empty(n, oh, ow, C.chunks_of(m.out_channels), 64, device=xs.device, dtype=torch.bfloat16) if split else None
        out = C.conv2d(xs, w, k, k, stride=s, padding=p, no_f32=not f32, cin=m.in_channels, beta=b, act=act, out_split=out_split)
        return out, out_split

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('S3FD runs on CUDA tensors only (no CPU fallback); got %s' % x.device)
        xs = C.pack_split32(x.float().permute(0, 2, 3, 1).contiguous())
        taps = {}
        trunk = list(_TRUNK)
        for i, spec in enumerate(trunk):
            if spec == '|':
                continue
            name = spec[0]
            pool_next = i + 1 < len(trunk) and trunk[i + 1] == '|'
            last = i + 1 == len(trunk)
            need_f32 = pool_next or name in _TAPS
            h, hs = self._conv(name, xs, 1, split=not (pool_next or last), f32=need_f32)
            if name in _TAPS:
                taps[name] = h
            if pool_next:                       # F.max_pool2d(h, 2, 2) on the channels-last tensor, then the next operand
                h = F.max_pool2d(h.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
                xs = C.pack_split32(h.contiguous())
            else:
                xs = hs
        outs = []
        for src, _, _ in _HEADS:
            if src.endswith('_norm'):
                f = taps[src[:-5]]
                ln = getattr(self, src)
                f = f / (f.pow(2).sum(dim=3, keepdim=True).sqrt() + ln.eps) * ln.weight          # net_s3fd.py:15-18, NHWC
            else:
                f = taps[src]
            fs = C.pack_split32(f.contiguous())
            cls, _ = self._conv(src + '_mbox_conf', fs, 0, split=False, f32=True)
            reg, _ = self._conv(src + '_mbox_loc', fs, 0, split=False, f32=True)
            outs += [cls.permute(0, 3, 1, 2), reg.permute(0, 3, 1, 2)]
        chunk = torch.chunk(outs[0], 4, 1)                                                       # max-out background label, :123-125
        outs[0] = torch.cat([torch.max(torch.max(chunk[0], chunk[1]), chunk[2]), chunk[3]], dim=1)
        return outs


def decode(loc, priors, variances=(0.1, 0.2)):
    """bbox.py:94-111: centre-offset predictions -> corner boxes."""
    boxes = torch.cat((priors[:, :2] + loc[:, :2] * variances[0] * priors[:, 2:],
                       priors[:, 2:] * torch.exp(loc[:, 2:] * variances[1])), 1)
    boxes[:, :2] -= boxes[:, 2:] / 2
    boxes[:, 2:] += boxes[:, :2]
    return boxes


def nms(dets, thresh):
    """bbox.py:48-66 on a [M, 5] numpy array -> indices kept."""
    if 0 == len(dets):
        return []
    x1, y1, x2, y2, scores = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        xx1, yy1 = np.maximum(x1[i], x1[order[1:]]), np.maximum(y1[i], y1[order[1:]])
        xx2, yy2 = np.minimum(x2[i], x2[order[1:]]), np.minimum(y2[i], y2[order[1:]])
        w, h = np.maximum(0.0, xx2 - xx1 + 1), np.maximum(0.0, yy2 - yy1 + 1)
        ovr = w * h / (areas[i] + areas[order[1:]] - w * h)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


def candidates(olist, j, conf=0.05):
    """Image j's candidate boxes [M, 5] (x1, y1, x2, y2, score) from the twelve head outputs (detect.py:38-58: soft-max, every
    position with face score > conf, box decoded from the anchor of its feature level: stride 4 .. 128, anchor 4 x stride).
    The reference collects the positions over the WHOLE batch for every image (a quirk of its np.where); the extra entries have
    scores at or below other images' and never survive its own NMS + 0.5 filter, so the final detections are the same."""
    rows = []
    for i in range(len(olist) // 2):
        ocls = F.softmax(olist[i * 2][j: j + 1].float(), dim=1)[0, 1]
        oreg = olist[i * 2 + 1][j].float()
        stride = 2 ** (i + 2)
        hh, ww = torch.nonzero(ocls > conf, as_tuple=True)
        if hh.numel() == 0:
            continue
        pri = torch.stack([stride / 2 + ww.float() * stride, stride / 2 + hh.float() * stride,
                           torch.full_like(ww, stride * 4, dtype=torch.float32), torch.full_like(ww, stride * 4, dtype=torch.float32)], 1)
        box = decode(oreg[:, hh, ww].t().contiguous(), pri)
        rows.append(torch.cat([box, ocls[hh, ww].unsqueeze(1)], 1))
    return torch.cat(rows).cpu().numpy() if rows else np.zeros((0, 5), dtype=np.float32)


def candidates_batch(olist, conf=0.05):
    """`candidates` for every image of the batch at once: per feature level one soft-max / threshold / decode over [B, H, W],
    one stable sort by image to restore the per-image, level-major order, ONE host synchronisation for the counts.
    Returns a list of B device tensors [M_j, 5]."""
    B = olist[0].shape[0]
    rows, owner = [], []
    for i in range(len(olist) // 2):
        ocls = F.softmax(olist[i * 2].float(), dim=1)[:, 1]                        # [B, H, W]
        oreg = olist[i * 2 + 1].float()                                            # [B, 4, H, W]
        stride = 2 ** (i + 2)
        bb, hh, ww = torch.nonzero(ocls > conf, as_tuple=True)
        if bb.numel() == 0:
            continue
        size = torch.full_like(ww, stride * 4, dtype=torch.float32)
        pri = torch.stack([stride / 2 + ww.float() * stride, stride / 2 + hh.float() * stride, size, size], 1)
        box = decode(oreg[bb, :, hh, ww].contiguous(), pri)
        rows.append(torch.cat([box, ocls[bb, hh, ww].unsqueeze(1)], 1))
        owner.append(bb)
    if not rows:
        return [olist[0].new_zeros(0, 5) for _ in range(B)]
    rows, owner = torch.cat(rows), torch.cat(owner)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=B).tolist()
    return list(torch.split(rows[order], counts))


def nms_device(dets, thresh):
    """bbox.py:48-66 on a device tensor [M, 5] -> kept indices in decreasing score order.  torchvision's CUDA NMS measures boxes
    as (x2 - x1) * (y2 - y1); the reference counts pixels, (x2 - x1 + 1) * (y2 - y1 + 1), in areas and intersections alike, which
    is the same thing on boxes whose far corner is moved out by one.  Falls back to the host loop above when torchvision's
    compiled ops are missing (the reference's own NMS is a host loop)."""
    try:
        from torchvision.ops import nms as tv_nms
        boxes = dets[:, :4].clone()
        boxes[:, 2:] += 1.0
        return tv_nms(boxes, dets[:, 4].contiguous(), thresh)
    except (ImportError, RuntimeError, NotImplementedError):
        keep = nms(dets.cpu().numpy(), thresh)
        return torch.as_tensor(np.asarray(keep, dtype=np.int64), device=dets.device)


class SFDDetector:
    """sfd_detector.SFDDetector: detect_from_batch(tensor [B, 3, H, W], RGB 0..255 - the traversal script feeds it without
    the BGR mean subtraction of the single-image path, sfd_detector.py:23-24) -> (per-image lists of [x1, y1, x2, y2, score]
    after NMS 0.3 and score > 0.5, error flag, index of the last image without a detection)."""

    def __init__(self, path_to_detector=None, device='cuda'):
        self.face_detector = S3FD()
        if path_to_detector is not None:
            self.face_detector.load_state_dict(torch.load(path_to_detector, map_location='cpu'))
        self.face_detector.to(device)

    @torch.no_grad()
    def detect_from_batch(self, tensor):
        olist = self.face_detector(tensor)
        out, error, error_index = [], False, -1
        for j, cand in enumerate(candidates_batch(olist)):
            if cand.shape[0] > 0:
                kept = cand[nms_device(cand, 0.3)]
                out.append(list(kept[kept[:, 4] > 0.5].cpu().numpy()))
            else:
                error, error_index = True, j
                out.append([])
        return out, error, error_index

    def __call__(self, tensor):
        """The `face_detector` callable of attribute_space.path_attributes."""
        return self.detect_from_batch(tensor)[0]
