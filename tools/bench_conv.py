"""Times the tensor-core conv at the StyleGAN2-1024 layer shapes and the full generator forward."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpedganspace_b200 import conv as C, _lib

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shapes = [(512, 512, 4), (512, 512, 8), (512, 512, 16), (512, 512, 32), (512, 512, 64), (256, 256, 128),
          (128, 128, 256), (64, 64, 512), (32, 32, 1024)]
print('same-res 3x3 convs, batch', B)
for ci, co, r in shapes:
    x = torch.randn(B, r, r, ci, device='cuda')
    w = torch.randn(co, ci, 3, 3, device='cuda') / (ci * 9) ** 0.5
    xs, ws = C.pack_split32(x), C.pack_weights(w)
    out = torch.empty(B, r, r, co, device='cuda')
    for bn in (0, 64, 128, 256):
        if bn > co: continue
        ms = timeit(lambda: C.conv2d(xs, ws, 3, 3, padding=1, out=out, force_bn=bn))
        fl = 2 * B * r * r * ci * co * 9
        print('  %4d->%4d @%4d bn=%3d  %8.3f ms  %7.1f TFLOP/s (x3 issued: %7.1f)' % (ci, co, r, bn, ms, fl / ms / 1e9, 3 * fl / ms / 1e9))
    tp = timeit(lambda: C.pack_split32(x, out=xs))
    print('     pack %.3f ms  (%.0f GB/s)' % (tp, 2 * x.numel() * 4 / tp / 1e6))
    del x, xs, out

if len(sys.argv) > 2:
    import oracle.stylegan2 as o
    from warpedganspace_b200.stylegan2 import Generator
    size = int(sys.argv[2])
    sd = o.init_state(size=size, generator=torch.Generator().manual_seed(0))
    G = Generator(size, 512, 8); G.load_state_dict(sd, strict=False); G.cuda().eval()
    z = torch.randn(B, 512, device='cuda')
    with torch.no_grad():
        ms = timeit(lambda: G([z]), iters=5, warm=2)
    print('G%d forward batch %d: %.2f ms  -> %.1f img/s, %.1f TFLOP/s algorithmic' % (size, B, ms, B / ms * 1e3, 148.5e9 * B / ms / 1e9 if size == 1024 else 0))
    _lib.reset_launch_count()
    with torch.no_grad(): G([z])
    print('launches per forward', _lib.launch_count())
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        with torch.no_grad(): G([z])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=12, max_name_column_width=60))
