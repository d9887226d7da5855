"""CPU: the phase-packed formulation of strided data-gradients / transposed convs (warpedganspace_b200.conv._phase_plan)
against torch, through a dense emulation of what the merged launch computes: every output phase (py, px) reads the same
small set of input shifts, the merged weight holds block (shift, phase) = tap or zero, and group g of the N columns lands
on output pixel (s*y + g // s, s*x + g % s).  The GPU tests check the kernel; this pins the index algebra."""
import pytest
import torch
import torch.nn.functional as F

from warpedganspace_b200.conv import _phase_plan


def _emulate(kind, x, w_src, kh, kw, s, p, out_hw):
    """x [N, H, W, K]; w_src [rows, K, T] -> out [N, oh, ow, rows] exactly as the merged conv launch lays it out."""
    shifts, idx, G = _phase_plan(kind, kh, kw, s, p, 'cpu')
    rows, K, T = w_src.shape
    w_ext = torch.cat([w_src, w_src.new_zeros(rows, K, 1)], 2)
    sel = w_ext.index_select(2, idx).permute(2, 0, 1).reshape(len(shifts), G * rows, K)
    N, H, W, _ = x.shape
    oh, ow = out_hw
    out = torch.zeros(N, oh, ow, rows, dtype=x.dtype)
    for q in range((oh + s - 1) // s):
        for r in range((ow + s - 1) // s):
            acc = torch.zeros(N, G * rows, dtype=x.dtype)
            for i, (sy, sx) in enumerate(shifts):
                iy, ix = q + sy, r + sx
                if 0 <= iy < H and 0 <= ix < W:
                    acc += x[:, iy, ix, :] @ sel[i].t()
            for g in range(G):
                Y, X = s * q + g // s, s * r + g % s
                if Y < oh and X < ow:
                    out[:, Y, X, :] = acc[:, g * rows:(g + 1) * rows]
    return out, shifts


@pytest.mark.parametrize('k,s,p,H,W', [(7, 2, 3, 12, 10), (3, 2, 1, 9, 8), (1, 2, 0, 8, 8), (3, 2, 1, 8, 7), (5, 2, 2, 11, 9),
                                       (3, 3, 1, 10, 10), (4, 2, 1, 8, 8)])
def test_strided_dgrad_as_one_merged_conv(k, s, p, H, W):
    torch.manual_seed(k * 100 + s * 10 + p)
    ci, co = 3, 5
    x = torch.randn(2, ci, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(co, ci, k, k, dtype=torch.float64)
    y = F.conv2d(x, w, stride=s, padding=p)
    dy = torch.randn_like(y)
    y.backward(dy)
    got, shifts = _emulate('dgrad', dy.permute(0, 2, 3, 1), w.permute(1, 0, 2, 3).reshape(ci, co, k * k), k, k, s, p, (H, W))
    assert float((got - x.grad.permute(0, 2, 3, 1)).abs().max()) < 1e-12
    assert len(shifts) <= ((k + s - 1) // s + 1) ** 2                     # far fewer tap loads than k*k


def test_stem_and_upconv_plans_have_the_documented_sizes():
    shifts, idx, G = _phase_plan('dgrad', 7, 7, 2, 3, 'cpu')            # ResNet stem: 49 taps over 4 phases -> 16 shifts
    assert len(shifts) == 16 and G == 4 and int((idx < 49).sum()) == 49
    shifts, idx, G = _phase_plan('convT', 3, 3, 2, 0, 'cpu')            # StyleGAN2 up-conv: 9 taps -> 4 shifts
    assert len(shifts) == 4 and G == 4 and int((idx < 9).sum()) == 9


@pytest.mark.parametrize('H,W', [(5, 6), (4, 4), (1, 3)])
def test_transposed_conv_as_one_merged_conv(H, W):
    torch.manual_seed(H * 10 + W)
    x = torch.randn(2, 4, H, W, dtype=torch.float64)
    w = torch.randn(3, 4, 3, 3, dtype=torch.float64)                                   # StyleGAN2 layout [Co, Ci, 3, 3]
    ref = F.conv_transpose2d(x, w.transpose(0, 1), stride=2)                           # models/StyleGAN2/model.py:206-212
    got, _ = _emulate('convT', x.permute(0, 2, 3, 1), w.reshape(3, 4, 9), 3, 3, 2, 0, (ref.shape[2], ref.shape[3]))
    assert float((got - ref.permute(0, 2, 3, 1)).abs().max()) < 1e-12
