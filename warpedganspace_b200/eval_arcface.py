"""ArcFace identity comparator of the attribute-space traversal (reference: lib/evaluation/archface/arcface.py:9-24 IDComparator,
:119-164 SE_IR(50, mode='ir_se')) on libwgs_b200.

IR-SE-50 is 24 residual units  BN -> conv3x3 -> PReLU -> conv3x3 (stride) -> BN -> squeeze-excite,  + shortcut (sub-sampling or
conv1x1 + BN), no activation after the sum.  Per unit here: the leading BatchNorm runs inside the operand pack (it cannot be
folded into a zero-padded conv), the trailing one is folded into the second conv, all three convs are tensor-core launches; the
PReLU, the squeeze-excite gate (two [N, C] x [C, C/16] products) and the final Linear + BatchNorm1d + L2 normalisation are torch
library calls on small tensors.  Parameter names are the reference's (``input_layer.0.weight`` ... ``body.23.res_layer.5.fc2.weight``,
``output_layer.4.running_var``), so ``model_ir_se50.pth`` loads as it is.  CUDA only.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import conv as C
from .eval_common import PackedConv, bn_affine, fold_bn, need_cuda
from .generators import affine_act_pack

_STAGES = ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3))      # arcface.py:103-108 (num_layers = 50)


class _SE(nn.Module):
    def __init__(self, c, reduction=16):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c // reduction, 1, bias=False)
        self.fc2 = nn.Conv2d(c // reduction, c, 1, bias=False)


class _Unit(nn.Module):
    """bottleneck_IR_SE (arcface.py:83-100) as a parameter container with the reference's sub-module indices."""

    def __init__(self, cin, depth, stride):
        super().__init__()
        self.stride = stride
        if cin == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(cin, depth, 1, stride, bias=False), nn.BatchNorm2d(depth))
        self.res_layer = nn.Sequential(nn.BatchNorm2d(cin), nn.Conv2d(cin, depth, 3, 1, 1, bias=False), nn.PReLU(depth),
                                       nn.Conv2d(depth, depth, 3, stride, 1, bias=False), nn.BatchNorm2d(depth), _SE(depth))


class IRSE50(nn.Module):
    """SE_IR(50, drop_ratio=0.4, mode='ir_se').forward in eval mode: x [N, 3, 112, 112] -> unit-norm embeddings [N, 512]."""

    def __init__(self):
        super().__init__()
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        self.output_layer = nn.Sequential(nn.BatchNorm2d(512), nn.Dropout(0.4), nn.Flatten(), nn.Linear(512 * 7 * 7, 512),
                                          nn.BatchNorm1d(512))
        self.body = nn.Sequential(*[_Unit(ci if i == 0 else d, d, 2 if i == 0 else 1) for ci, d, n in _STAGES for i in range(n)])
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
        if self._plan is not None:
            return self._plan
        P = {}
        with torch.no_grad():
            il = self.input_layer
            P['stem'] = (PackedConv(*fold_bn(il[0].weight, None, il[1]), 1, 1), il[2].weight.float())
            P['units'] = []
            for u in self.body:
                r = u.res_layer
                e = dict(stride=u.stride, pre=bn_affine(r[0]), conv1=PackedConv(r[1].weight, None, 1, 1), slope=r[2].weight.float(),
                         conv2=PackedConv(*fold_bn(r[3].weight, None, r[4]), u.stride, 1),
                         fc1=r[5].fc1.weight.float().flatten(1).t().contiguous(), fc2=r[5].fc2.weight.float().flatten(1).t().contiguous())
                if isinstance(u.shortcut_layer, nn.Sequential):
                    e['shortcut'] = PackedConv(*fold_bn(u.shortcut_layer[0].weight, None, u.shortcut_layer[1]), u.stride, 0)
                P['units'].append(e)
            ol = self.output_layer
            a, b = bn_affine(ol[0])
            # BatchNorm2d -> Flatten (NCHW order) -> Linear, on the NHWC map: the affine goes into the Linear's columns
            w = ol[3].weight.float().view(512, 512, 7, 7).permute(0, 2, 3, 1)                       # [out, h, w, c]
            P['fc_w'] = (w * a.view(1, 1, 1, -1)).reshape(512, -1).contiguous()
            P['fc_b'] = (ol[3].bias.float() + (w * b.view(1, 1, 1, -1)).sum(dim=(1, 2, 3))).contiguous()
            s = ol[4].weight / torch.sqrt(ol[4].running_var + ol[4].eps)
            P['bn1d'] = (s.float(), (ol[4].bias - ol[4].running_mean * s).float())
        self._plan = P
        return P

    def forward(self, x):
        need_cuda(x, 'IRSE50')
        P = self.plan()
        conv, slope = P['stem']
        y, _ = conv(C.pack_split32(x.float().permute(0, 2, 3, 1).contiguous()))
        cur = torch.where(y >= 0, y, y * slope)
        for e in P['units']:
            s = e['stride']
            y, _ = e['conv1'](affine_act_pack(cur, e['pre'][0], e['pre'][1], relu=False))
            res, _ = e['conv2'](C.pack_split32(torch.where(y >= 0, y, y * e['slope'])))
            gate = torch.sigmoid(F.relu(res.mean(dim=(1, 2)) @ e['fc1']) @ e['fc2'])              # SEModule, arcface.py:51-58
            if 'shortcut' in e:
                sc, _ = e['shortcut'](C.pack_split32(cur))
            else:
                sc = cur[:, ::s, ::s]                                                              # MaxPool2d(1, stride)
            cur = torch.addcmul(sc, res, gate[:, None, None, :]).contiguous()
        f = torch.addmm(P['fc_b'], cur.flatten(1), P['fc_w'].t())
        f = f * P['bn1d'][0] + P['bn1d'][1]
        return f / torch.norm(f, 2, 1, True)


class IDComparator(nn.Module):
    """arcface.IDComparator: forward(x, x_prime) -> mean cosine similarity of the embeddings of the fixed face region
    [35:223, 32:220] of two batches of 256 x 256 frames in -1 .. 1, pooled to 112 x 112."""

    def __init__(self, path_to_backbone=None):
        super().__init__()
        self.backbone = IRSE50()
        if path_to_backbone is not None:
            self.backbone.load_state_dict(torch.load(path_to_backbone, map_location='cpu'))
        self.face_pool = nn.AdaptiveAvgPool2d((112, 112))
        self.criterion = nn.CosineSimilarity(dim=1, eps=1e-6)

    def extract_feats(self, x):
        return self.backbone(self.face_pool(x[:, :, 35:223, 32:220]))

    def forward(self, x, x_prime):
        return self.criterion(self.extract_feats(x), self.extract_feats(x_prime)).mean()
