"""CPU study (no GPU): image error of StyleGAN2 when the two correction products of the fp32-accurate split
(a_hi*b_lo, a_lo*b_hi) are carried in fp8 instead of bf16 - the option DESIGN.md section 11 item 4 names.

Every modulated conv of the oracle generator is replaced by  d * conv(split(W*scale), split(s*x))  with the operand
rounding of one scheme and fp64 accumulation; the image is compared with the all-fp64 forward.  Schemes:
  bf16x3        a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, all bf16                       (what libwgs_b200 issues today)
  bf16x1        a_hi*b_hi only
  tf32          one product of operands rounded to 10 mantissa bits
  lo8           a_hi*b_hi + a_hi*b_lo in bf16, a_lo*b_hi with BOTH factors in e4m3 (per-layer power-of-two scales,
                separate accumulator)                                            -> A bytes 128 -> 96 per pixel-chunk
  corr8         a_hi*b_hi in bf16, both correction products in e4m3
  corr8_e5m2b   as corr8 with the weight-side factors in e5m2
Usage: python tools/precision_study.py [size=128] [batch=2] [scheme,scheme,...]
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import oracle.stylegan2 as o


def bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def split(x):
    hi = bf16(x)
    return hi, bf16(x - hi)


def tf32(x):
    x32 = x.to(torch.float32)
    i = x32.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF                      # round to nearest, 10 explicit mantissa bits
    return i.view(torch.float32).to(torch.float64)


def q8(x, dtype):
    """fp8 with a per-tensor power-of-two scale chosen so that max|x| lands just below the format's top binade."""
    top = 256.0 if dtype == torch.float8_e4m3fn else 32768.0
    m = float(x.abs().max())
    if m == 0.0:
        return x
    s = 2.0 ** math.floor(math.log2(top / m))
    return (x * s).to(torch.float32).to(dtype).to(torch.float64) / s


def products(scheme, a, b):
    """list of (a_term, b_term) whose convs are summed."""
    if scheme == 'fp64':
        return [(a, b)]
    if scheme == 'tf32':
        return [(tf32(a), tf32(b))]
    ah, al = split(a)
    bh, bl = split(b)
    if scheme == 'bf16x1':
        return [(ah, bh)]
    if scheme == 'bf16x3':
        return [(ah, bh), (ah, bl), (al, bh)]
    e4, e5 = torch.float8_e4m3fn, torch.float8_e5m2
    if scheme == 'lo8':
        return [(ah, bh), (ah, bl), (q8(al, e4), q8(bh, e4))]
    if scheme == 'corr8':
        return [(ah, bh), (q8(ah, e4), q8(bl, e4)), (q8(al, e4), q8(bh, e4))]
    if scheme == 'corr8_e5m2b':
        return [(ah, bh), (q8(ah, e4), q8(bl, e5)), (q8(al, e4), q8(bh, e5))]
    raise ValueError(scheme)


def make_modulated_conv(scheme):
    def modulated_conv(sd, prefix, x, w, demodulate=True, upsample=False, blur_taps=(1, 3, 3, 1)):
        weight = sd[prefix + '.weight'][0].double()                                # [Co, Ci, k, k]
        co, ci, k, _ = weight.shape
        style = o.equal_linear(w, sd[prefix + '.modulation.weight'], sd[prefix + '.modulation.bias']).double()
        scale = 1.0 / math.sqrt(ci * k * k)
        wq = weight * scale
        a = x.double() * style[:, :, None, None]                                   # activation carries the style
        y = 0
        for at, bt in products(scheme, a, wq):
            if upsample:
                y = y + F.conv_transpose2d(at, bt.transpose(0, 1), stride=2)
            else:
                y = y + F.conv2d(at, bt, padding=k // 2)
        if demodulate:
            d = torch.rsqrt((scale ** 2) * (style.pow(2) @ weight.pow(2).sum([2, 3]).t()) + 1e-8)
            y = y * d[:, :, None, None]
        if upsample:
            p = (len(blur_taps) - 2) - (k - 1)
            kern = o.fir_kernel(blur_taps).double() * 4.0
            y = o.upfirdn2d(y, kern, pad=((p + 1) // 2 + 1, p // 2 + 1))
        return y
    return modulated_conv


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(0)
    sd = o.init_state(size=size, generator=g)
    for k_ in list(sd):
        if k_.endswith('noise.weight'):
            sd[k_] = torch.full_like(sd[k_], 0.1)
    sd = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}
    z = torch.randn(batch, 512, generator=g).double()
    real = o.modulated_conv
    images = {}
    schemes = sys.argv[3].split(',') if len(sys.argv) > 3 else ['bf16x3', 'lo8', 'corr8', 'corr8_e5m2b', 'tf32', 'bf16x1']
    for scheme in ['fp64'] + schemes:
        o.modulated_conv = make_modulated_conv(scheme)
        with torch.no_grad():
            images[scheme] = o.generate(sd, z, None, size)
        o.modulated_conv = real
        if scheme != 'fp64':
            ref = images['fp64']
            e = images[scheme] - ref
            print('| %-12s | %.2e | %.2e |' % (scheme, float(e.norm() / ref.norm()), float(e.abs().max() / ref.abs().max())),
                  flush=True)


if __name__ == '__main__':
    main()
