"""Times the hot StyleGAN2-1024 layer shapes through the C ABI with CUDA events (L2 flushed between launches by
cycling through more than 126 MB of operands) - plain and fused epilogues.  Usage: python tools/bench_layers.py [which]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpedganspace_b200 import conv as C


def timeit(fn, iters=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def layer(ci, co, r, B, fused, taps=3):
    x = torch.randn(B, r, r, ci, device='cuda')
    w = torch.randn(co, ci, taps, taps, device='cuda') / (ci * taps * taps) ** 0.5
    xs, ws = C.pack_split32(x), C.pack_weights(w)
    del x
    alpha = torch.rand(B, co, device='cuda') + 0.5
    beta = torch.randn(co, device='cuda')
    noise = torch.randn(r, r, device='cuda')
    if not fused:
        out = torch.empty(B, r, r, co, device='cuda')
        fn = lambda: C.conv2d(xs, ws, taps, taps, padding=taps // 2, out=out, alpha=alpha, beta=beta, noise=noise, noise_w=0.1, act=3)
    else:
        out = torch.empty(B, r, r, co, device='cuda')
        nxt = torch.empty(B, r, r, co // 32, 64, dtype=torch.bfloat16, device='cuda')
        sc = torch.rand(B, co, device='cuda')
        rgb_w = torch.randn(B, 3, co, device='cuda')
        rgb = torch.zeros(B, r, r, 3, device='cuda')
        fn = lambda: C.conv2d(xs, ws, taps, taps, padding=taps // 2, out=out, alpha=alpha, beta=beta, noise=noise, noise_w=0.1,
                              act=3, out_split=nxt, split_scale=sc, out_from_n=B // 2, rgb_w=rgb_w, rgb_out=rgb)
    ms = timeit(fn)
    fl = 2.0 * B * r * r * ci * co * taps * taps
    print('%4d->%4d @%4d B=%d %s: %.3f ms  %.1f TFLOP/s algorithmic' % (ci, co, r, B, 'fused' if fused else 'plain', ms, fl / ms / 1e9), flush=True)


def variants(ci, co, r, B):
    """Which part of the fused epilogue costs what (32 -> 32 @1024^2)."""
    x = torch.randn(B, r, r, ci, device='cuda')
    w = torch.randn(co, ci, 3, 3, device='cuda') / (ci * 9) ** 0.5
    xs, ws = C.pack_split32(x), C.pack_weights(w)
    del x
    alpha = torch.rand(B, co, device='cuda') + 0.5
    beta = torch.randn(co, device='cuda')
    noise = torch.randn(r, r, device='cuda')
    out = torch.empty(B, r, r, co, device='cuda')
    nxt = torch.empty(B, r, r, co // 32, 64, dtype=torch.bfloat16, device='cuda')
    sc = torch.rand(B, co, device='cuda')
    rgb_w = torch.randn(B, 3, co, device='cuda')
    rgb = torch.zeros(B, r, r, 3, device='cuda')
    base = dict(padding=1, alpha=alpha, beta=beta, noise=noise, noise_w=0.1, act=3)
    cases = {
        'plain fp32 all': dict(out=out),
        'plain no noise/alpha/beta': dict(out=out, alpha=None, beta=None, noise=None, noise_w=0.0, act=0),
        'fused fp32 all (from_n=0 + dummy split off)': dict(out=out, out_from_n=1),
        'fused fp32 half': dict(out=out, out_from_n=B // 2),
        'fused split only': dict(out=None, no_f32=True, out_split=nxt, split_scale=sc),
        'fused rgb only': dict(out=None, no_f32=True, rgb_w=rgb_w, rgb_out=rgb),
        'fused split+rgb': dict(out=None, no_f32=True, out_split=nxt, split_scale=sc, rgb_w=rgb_w, rgb_out=rgb),
        'fused split+fp32 half': dict(out=out, out_split=nxt, split_scale=sc, out_from_n=B // 2),
        'fused all': dict(out=out, out_split=nxt, split_scale=sc, out_from_n=B // 2, rgb_w=rgb_w, rgb_out=rgb),
    }
    for name, kw in cases.items():
        args = dict(base); args.update(kw)
        ms = timeit(lambda: C.conv2d(xs, ws, 3, 3, **args))
        print('  %-46s %.3f ms' % (name, ms), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', '32'):
    layer(32, 32, 1024, 8, False)
    layer(32, 32, 1024, 8, True)
    layer(32, 32, 1024, 4, False)
if which in ('all', '64'):
    layer(64, 64, 512, 8, False)
    layer(64, 64, 512, 8, True)
if which == 'var':
    variants(32, 32, 1024, 8)
if which == 'all':
    layer(128, 128, 256, 8, True)
    layer(512, 512, 64, 8, True)
