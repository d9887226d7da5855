"""Action-unit detector of the attribute-space traversal (reference: lib/evaluation/au_detector/hourglass.py:17-243 FANAU over
QFAN, AU_detector.py:29-46 AUdetector.detect_AU) on libwgs_b200.

A face-alignment network (7x7 / 2 stem, three ConvBlocks, one depth-4 hourglass, 68 landmark heat maps) feeds a second,
lightweight depth-4 hourglass that regresses 12 action-unit heat maps; the intensity of an action unit is the maximum of its
64 x 64 map.  Every convolution (about 110 of them) is one tensor-core launch with its eval-mode BatchNorm folded into the
weights and the bias + ReLU in the epilogue; the ReLU6 clamp, the channel concatenation + residual sum of a ConvBlock, the 2x2
max-pools and the nearest-neighbour up-sampling of the hourglass are torch library calls on <= 64 x 64 maps.  Parameter names are
the reference's (``fan.m0.b2_plus_1.conv3.weight`` ...), so ``disfa_adaptation_f0.pth``'s ``state_dict`` loads as it is.  CUDA only.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import conv as C
from .eval_common import PackedConv, fold_bn, need_cuda


class _ConvBlock(nn.Module):
    """hourglass.py:17-68 as a parameter container: three convs (3x3, or 1x1 when lightweight) with BatchNorm whose outputs
    are concatenated (out/2 + out/4 + out/4 channels) and added to the (1x1-projected when in != out) input."""

    def __init__(self, cin, cout, lightweight=False):
        super().__init__()
        k = 1 if lightweight else 3
        chans = ((cin, cout // 2), (cout // 2, cout // 4), (cout // 4, cout // 4))
        for j, (a, b) in enumerate(chans, start=1):
            setattr(self, 'conv%d' % j, nn.Conv2d(a, b, k, 1, (k - 1) // 2, bias=False))
        for j, (a, b) in enumerate(chans, start=1):
            setattr(self, 'bn%d' % j, nn.BatchNorm2d(b))
        self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU6(True)) \
            if cin != cout else None


class _HourGlass(nn.Module):
    def __init__(self, depth, features, lightweight=False):
        super().__init__()
        self.depth = depth
        self._gen(depth, features, lightweight)

    def _gen(self, level, f, lightweight):                     # registration order of hourglass.py:78-89
        self.add_module('b1_%d' % level, _ConvBlock(f, f, lightweight))
        self.add_module('b2_%d' % level, _ConvBlock(f, f))
        if level > 1:
            self._gen(level - 1, f, lightweight)
        else:
            self.add_module('b2_plus_%d' % level, _ConvBlock(f, f))
        self.add_module('b3_%d' % level, _ConvBlock(f, f))


class _QFAN(nn.Module):
    def __init__(self, f=128, num_out=68):
        super().__init__()
        self.conv1 = nn.Conv2d(3, f // 2, 7, 2, 3)
        self.bn1 = nn.BatchNorm2d(f // 2)
        self.conv2, self.conv3, self.conv4 = _ConvBlock(f // 2, f // 2), _ConvBlock(f // 2, f), _ConvBlock(f, f)
        self.m0 = _HourGlass(4, f)
        self.top_m_0 = _ConvBlock(f, f)
        self.conv_last0 = nn.Conv2d(f, f, 1)
        self.bn_end0 = nn.BatchNorm2d(f)
        self.l0 = nn.Conv2d(f, num_out, 1)


def _relu6_(t):
    return t.clamp_(max=6.0)                                   # on top of the epilogue's ReLU


class FANAU(nn.Module):
    """hourglass.FANAU(num_modules=1, n_points=12).forward in eval mode: x [N, 3, 256, 256] in 0..1 -> heat maps [N, 12, 64, 64]."""

    def __init__(self, n_points=12, f=128):
        super().__init__()
        self.fan = _QFAN(f)
        self.conv1 = nn.Sequential(nn.Conv2d(68, f, 1, 1), nn.BatchNorm2d(f), nn.ReLU6())
        self.conv2 = nn.Sequential(nn.Conv2d(f, f, 1, 1), nn.BatchNorm2d(f), nn.ReLU6())
        self.net = _HourGlass(4, f, lightweight=True)
        self.conv_last = nn.Sequential(nn.Conv2d(f, f, 1, 1), nn.BatchNorm2d(f), nn.ReLU6())
        self.l = nn.Conv2d(f, n_points, 1, 1)
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
    def plan(self):
        if self._plan is not None:
            return self._plan
        P = {}
        with torch.no_grad():
            for name, m in self.named_modules():
                if isinstance(m, _ConvBlock):
                    e = [PackedConv(*fold_bn(getattr(m, 'conv%d' % j).weight, None, getattr(m, 'bn%d' % j)), 1,
                                    getattr(m, 'conv%d' % j).padding[0]) for j in (1, 2, 3)]
                    down = PackedConv(*fold_bn(m.downsample[0].weight, None, m.downsample[1]), 1, 0) if m.downsample is not None else None
                    P[name] = (e, down)
            w, b = fold_bn(self.fan.conv1.weight, self.fan.conv1.bias, self.fan.bn1)
            P['stem'] = (C.pack_weights_group([(w, C.PACK_S2D, 3)])[0], b)
            P['fan.conv_last0'] = PackedConv(*fold_bn(self.fan.conv_last0.weight, self.fan.conv_last0.bias, self.fan.bn_end0))
            P['fan.l0'] = PackedConv(self.fan.l0.weight, self.fan.l0.bias)
            for n in ('conv1', 'conv2', 'conv_last'):
                seq = getattr(self, n)
                P[n] = PackedConv(*fold_bn(seq[0].weight, seq[0].bias, seq[1]))
            P['l'] = PackedConv(self.l.weight, self.l.bias)
        self._plan = P
        return P

    def _block(self, name, x):
        (c1, c2, c3), down = self.plan()[name]
        xs = C.pack_split32(x)
        o1 = _relu6_(c1(xs, act=1)[0])
        o2 = _relu6_(c2(C.pack_split32(o1), act=1)[0])
        o3 = _relu6_(c3(C.pack_split32(o2), act=1)[0])
        res = _relu6_(down(xs, act=1)[0]) if down is not None else x
        return torch.cat((o1, o2, o3), dim=3).add_(res)

    def _hourglass(self, name, level, x):                      # hourglass.py:91-113 on NHWC maps
        up1 = self._block('%s.b1_%d' % (name, level), x)
        low = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()
        low = self._block('%s.b2_%d' % (name, level), low)
        low = self._hourglass(name, level - 1, low) if level > 1 else self._block('%s.b2_plus_%d' % (name, level), low)
        low = self._block('%s.b3_%d' % (name, level), low)
        return up1 + low.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)

    def forward(self, x):
        need_cuda(x, 'FANAU')
        P = self.plan()
        n, _, h, w = x.shape
        assert h % 4 == 0 and w % 4 == 0, 'FANAU needs frame sizes divisible by 4 (the reference feeds 256 x 256)'
        w_stem, b_stem = P['stem']
        taps, _ = C.s2d_taps(7, 3)
        y = torch.empty(n, h // 2, w // 2, 64, device=x.device, dtype=torch.float32)
        C.conv_taps(C.s2d_pack_split32(x.float().permute(0, 2, 3, 1).contiguous()), w_stem, taps, y, grid=(h // 2, w // 2),
                    cin=7 * 7 * 3, beta=b_stem, act=1, algo_macs_per_pixel=7 * 7 * 3 * 64)
        y = self._block('fan.conv2', y)
        y = F.max_pool2d(y.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()
        feat = self._block('fan.conv4', self._block('fan.conv3', y))
        ll = self._block('fan.top_m_0', self._hourglass('fan.m0', 4, feat))
        _, lls = P['fan.conv_last0'](C.pack_split32(ll), act=1, f32=False, split=True)
        lmk, _ = P['fan.l0'](lls)                                                         # 68 landmark heat maps
        a = _relu6_(P['conv1'](C.pack_split32(lmk), act=1)[0])
        b = _relu6_(P['conv2'](C.pack_split32(feat), act=1)[0])
        hm = self._hourglass('net', 4, a.add_(b))
        hm = _relu6_(P['conv_last'](C.pack_split32(hm), act=1)[0])
        out, _ = P['l'](C.pack_split32(hm))
        return out.permute(0, 3, 1, 2)


class AUdetector:
    """AU_detector.AUdetector: detect_AU(img [N, 3, 256, 256] or [3, 256, 256]) -> intensities [N, 12]."""

    def __init__(self, au_model_path=None, device='cuda'):
        self.naus = 12
        self.FAN = FANAU(n_points=self.naus)
        if au_model_path is not None:
            self.FAN.load_state_dict(torch.load(au_model_path, map_location='cpu')['state_dict'])
        self.FAN.to(device)

    @torch.no_grad()
    def detect_AU(self, img):
        x = (img - img.min()) / (img.max() - img.min())
        if x.ndim == 3:
            x = x.unsqueeze(0)
        heat = self.FAN(x.to(next(self.FAN.parameters()).device))
        return F.max_pool2d(heat, (64, 64)).squeeze(2).squeeze(2)

    def __call__(self, img):
        """The `au_detector` callable of attribute_space.path_attributes."""
        return self.detect_AU(img)
