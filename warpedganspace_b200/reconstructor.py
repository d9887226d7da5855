"""Reconstructor: LeNet / ResNet-18 regressor-classifier with its convolutions on libwgs_b200.

Drop-in for the reference ``lib.reconstructor.Reconstructor`` (lib/reconstructor.py:10-79): same
constructor signature, ``forward(x1, x2) -> (logits [B, dim], magnitudes [B])`` and the same state-dict
keys (ResNet: torchvision ``resnet18`` names under ``features_extractor.`` including the unused ``fc``;
LeNet: ``feature_extractor.{0,1,4,5,8,9}``, heads ``{path_indices,shift_magnitudes}.{0,1,3}``).

Every convolution (forward and data-gradient) runs on the tcgen05 tap-list kernel; activations stay NHWC
(torch ``channels_last``).  BatchNorm keeps the reference's train-mode batch statistics.
"""
import math
import os
import weakref

import torch
from torch import nn
import torch.nn.functional as F

from . import _lib
from . import conv as C
from . import wgrad as WG


_FROZEN_PACKS = {}
MERGE_PHASES = os.environ.get('WGS_MERGE_PHASES', '1') != '0'      # 0 = one launch per output phase (A/B switch)


def _packed(w, transposed):
    """split32 pack of a conv weight; cached for frozen (requires_grad=False) weights, which never change."""
    if w.requires_grad:
        return C.pack_weights(w.detach().permute(1, 0, 2, 3).contiguous() if transposed else w.detach())
    key = (w.data_ptr(), tuple(w.shape), w._version, transposed)
    hit = _FROZEN_PACKS.get(key)
    if hit is not None and hit[0]() is not w:      # the address was freed and reused by ANOTHER tensor: stale entry
        hit = None
    if hit is None:
        if len(_FROZEN_PACKS) > 512:
            _FROZEN_PACKS.clear()
        hit = (weakref.ref(w), C.pack_weights(w.detach().permute(1, 0, 2, 3).contiguous() if transposed else w.detach()))
        _FROZEN_PACKS[key] = hit
    return hit[1]


class _ConvFn(torch.autograd.Function):
    """F.conv2d(x, w, bias, stride, padding) on the tensor-core kernel; x, out are logical NCHW with
    channels-last memory."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, padding):
        if not x.is_cuda:
            raise RuntimeError('Reconstructor runs on CUDA tensors only (no CPU fallback); got %s' % x.device)
        co, ci, kh, kw = w.shape
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        xs = C.pack_split32(x_nhwc)
        ws = _packed(w, False)
        out = C.conv2d(xs, ws, kh, kw, stride=stride, padding=padding, cin=ci,
                       beta=bias.detach() if bias is not None else None)
        ctx.save_for_backward(xs, w)
        ctx.geom = (x.shape, stride, padding, bias is not None)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xs, w = ctx.saved_tensors
        (n, ci, h, wd), stride, padding, has_bias = ctx.geom
        co, _, kh, kw = w.shape
        dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
        dys = C.pack_split32(dy_nhwc)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad(dys, w, (h, wd), stride, padding).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dw = WG.conv_wgrad(xs, dys, (co, ci, kh, kw), stride, padding)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy_nhwc.sum(dim=(0, 1, 2))
        return dx, dw, db, None, None


class _StemConvFn(torch.autograd.Function):
    """Few-input-channel convolution (the 7x7/2 stem on 6 channels) as im2col -> one 1-tap GEMM on the tensor
    cores: K = kh*kw*Ci packed densely instead of kh*kw taps of a mostly-empty 32-channel chunk.  The data
    gradient (needed: it carries d loss / d image back into the generator) stays on the phase tap-list path."""

    @staticmethod
    def forward(ctx, x, w, stride, padding):
        if not x.is_cuda:
            raise RuntimeError('Reconstructor runs on CUDA tensors only (no CPU fallback); got %s' % x.device)
        co, ci, kh, kw = w.shape
        n, _, h, wd = x.shape
        oh = (h + 2 * padding - kh) // stride + 1
        ow = (wd + 2 * padding - kw) // stride + 1
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        chunks = (kh * kw * ci + 31) // 32
        xcol = torch.empty(n, oh, ow, chunks, 64, device=x.device, dtype=torch.bfloat16)
        _lib.call('wgs_im2col_split32', _lib.ptr(x_nhwc), n, h, wd, ci, kh, kw, stride, padding, oh, ow,
                  _lib.ptr(xcol), _lib.stream())
        wmat = w.detach().permute(0, 2, 3, 1).reshape(co, kh * kw * ci, 1, 1)        # K order (ky, kx, c)
        out = C.conv2d(xcol, C.pack_weights(wmat), 1, 1, cin=kh * kw * ci)
        ctx.save_for_backward(xcol, w)
        ctx.geom = (x.shape, stride, padding)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xcol, w = ctx.saved_tensors
        (n, ci, h, wd), stride, padding = ctx.geom
        co, _, kh, kw = w.shape
        dys = C.pack_split32(dy.permute(0, 2, 3, 1).contiguous())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad(dys, w, (h, wd), stride, padding).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dwm = WG.conv_wgrad(xcol, dys, (co, kh * kw * ci, 1, 1), 1, 0)           # [Co, K, 1, 1]
            dw = dwm.reshape(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
        return dx, dw, None, None


def conv_dgrad(dys, w, in_hw, stride, padding, out=None, accumulate=False, packed=None, split_k=False):
    """Data gradient of F.conv2d: dx[iy,ix,ci] = sum_{ky,kx,co} dy[(iy+p-ky)/s, (ix+p-kx)/s, co] * w[co,ci,ky,kx]
    (terms with non-integer quotients vanish).  stride 1: one flipped-tap conv; stride s: s*s output phases.
    packed: the weight already packed for this call (transposed taps for stride 1, phase-merged blocks for stride > 1;
    conv.pack_weights_group) - otherwise it is packed here."""
    co, ci, kh, kw = w.shape
    h, wd = in_hw
    n, oh, ow = dys.shape[0], dys.shape[1], dys.shape[2]
    if stride > 1 and MERGE_PHASES and stride * stride * ci <= 1024:
        # all stride*stride output phases in one launch (phase-packed output, conv.py)
        return C.conv_dgrad_merged(dys, w, in_hw, stride, padding, out=out, accumulate=accumulate, w_merged=packed)
    wt = packed if packed is not None else _packed(w, True)                 # [kh*kw, Ci, Co]
    dx = out if out is not None else torch.empty(n, h, wd, ci, device=dys.device, dtype=torch.float32)
    for py in range(stride):
        for px in range(stride):
            gh = (h - py + stride - 1) // stride
            gw = (wd - px + stride - 1) // stride
            if gh <= 0 or gw <= 0:
                continue
            kys = [ky for ky in range(kh) if (py + padding - ky) % stride == 0]
            kxs = [kx for kx in range(kw) if (px + padding - kx) % stride == 0]
            if not kys or not kxs:
                if not accumulate:
                    dx[:, py::stride, px::stride, :] = 0
                continue
            taps = [((py + padding - ky) // stride, (px + padding - kx) // stride, ky * kw + kx)
                    for ky in kys for kx in kxs]
            C.conv_taps(dys, wt, taps, dx, grid=(gh, gw), out_origin=(py, px), out_step=(stride, stride), cout=ci, cin=co,
                        accumulate=accumulate, split_k=split_k)
    return dx


def conv2d(x, w, bias=None, stride=1, padding=0):
    if bias is None and w.shape[1] < 16 and w.shape[2] * w.shape[3] >= 9:
        return _StemConvFn.apply(x, w, stride, padding)
    return _ConvFn.apply(x, w, bias, stride, padding)


class _Node(nn.Module):
    pass


def _conv_param(co, ci, k, bias=False, resnet_init=True):
    m = _Node()
    w = torch.empty(co, ci, k, k)
    if resnet_init:
        nn.init.kaiming_normal_(w, mode='fan_out', nonlinearity='relu')
    else:
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    m.weight = nn.Parameter(w)
    if bias:
        bound = 1.0 / math.sqrt(ci * k * k)
        m.bias = nn.Parameter(torch.empty(co).uniform_(-bound, bound))
    return m


class _ResNet18(nn.Module):
    """torchvision resnet18 parameter tree (names only; compute is in Reconstructor.forward)."""

    def __init__(self):
        super().__init__()
        self.conv1 = _conv_param(64, 6, 7)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (c, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            blocks = nn.ModuleList()
            for bi in range(2):
                b = _Node()
                b.conv1 = _conv_param(c, cin if bi == 0 else c, 3)
                b.bn1 = nn.BatchNorm2d(c)
                b.conv2 = _conv_param(c, c, 3)
                b.bn2 = nn.BatchNorm2d(c)
                b.stride = stride if bi == 0 else 1
                if bi == 0 and (stride != 1 or cin != c):
                    b.downsample = nn.ModuleList([_conv_param(c, cin, 1), nn.BatchNorm2d(c)])
                blocks.append(b)
            setattr(self, 'layer%d' % li, blocks)
            cin = c
        self.fc = nn.Linear(512, 1000)             # present in the reference state dict, never used
        for p in self.fc.parameters():
            p.requires_grad_(False)




class Reconstructor(nn.Module):
    def __init__(self, reconstructor_type, dim, channels=3):
        super().__init__()
        self.reconstructor_type = reconstructor_type
        self.dim = dim
        self.channels = channels
        self.fused = True            # ResNet trunk as one hand-scheduled autograd node (train mode); False = per-op autograd
        if reconstructor_type == 'LeNet':
            self.lenet_width = 2
            fe = nn.ModuleList([nn.Identity() for _ in range(11)])
            for idx, (ci, co) in zip((0, 4, 8), ((channels * 2, 6), (6, 16), (16, 120))):
                fe[idx] = _conv_param(co, ci, 5, bias=True, resnet_init=False)
                fe[idx + 1] = nn.BatchNorm2d(co)
            self.feature_extractor = fe
            self.path_indices = nn.Sequential(nn.Linear(120, 84), nn.BatchNorm1d(84), nn.ReLU(), nn.Linear(84, dim))
            self.shift_magnitudes = nn.Sequential(nn.Linear(120, 84), nn.BatchNorm1d(84), nn.ReLU(), nn.Linear(84, 1))
        elif reconstructor_type == 'ResNet':
            self.features_extractor = _ResNet18()
            self.path_indices = nn.Linear(512, dim)
            self.shift_magnitudes = nn.Linear(512, 1)
        else:
            raise ValueError('reconstructor_type must be LeNet or ResNet')

    def features(self, x):
        if self.reconstructor_type == 'LeNet':
            fe = self.feature_extractor
            x = F.max_pool2d(F.relu(fe[1](conv2d(x, fe[0].weight, fe[0].bias))), 2, 2)
            x = F.max_pool2d(F.relu(fe[5](conv2d(x, fe[4].weight, fe[4].bias))), 2, 2)
            x = F.relu(fe[9](conv2d(x, fe[8].weight, fe[8].bias)))
            return x.mean(dim=[-1, -2]).view(x.shape[0], -1)
        r = self.features_extractor
        if self.fused and self.training:
            from .resnet_fused import ResNetFeatures
            return ResNetFeatures.apply(x, None, r, *ResNetFeatures.param_list(r))
        x = F.relu(r.bn1(conv2d(x, r.conv1.weight, None, 2, 3)))
        x = F.max_pool2d(x, 3, 2, 1)
        for li in range(1, 5):
            for b in getattr(r, 'layer%d' % li):
                idt = x
                h = F.relu(b.bn1(conv2d(x, b.conv1.weight, None, b.stride, 1)))
                h = b.bn2(conv2d(h, b.conv2.weight, None, 1, 1))
                if hasattr(b, 'downsample'):
                    idt = b.downsample[1](conv2d(x, b.downsample[0].weight, None, b.stride, 0))
                x = F.relu(h + idt)
        return x.mean(dim=[2, 3])

    def forward(self, x1, x2):
        if self.reconstructor_type == 'ResNet' and self.fused and self.training and x1.shape == x2.shape:
            # torch.cat([x1, x2], dim=1) (lib/reconstructor.py:72) folded into the stem's operand pack
            from .resnet_fused import ResNetFeatures
            r = self.features_extractor
            f = ResNetFeatures.apply(x1, x2, r, *ResNetFeatures.param_list(r))
            return self.path_indices(f), self.shift_magnitudes(f).squeeze()
        x = torch.cat([x1, x2], dim=1).contiguous(memory_format=torch.channels_last)
        f = self.features(x)
        return self.path_indices(f), self.shift_magnitudes(f).squeeze()
