"""Launches the tensor-core conv at three representative StyleGAN2-1024 layer shapes (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpedganspace_b200 import conv as C
B = 8
for ci, co, r in [(512, 512, 64), (64, 64, 512), (32, 32, 1024)]:
    x = torch.randn(B, r, r, ci, device='cuda')
    w = torch.randn(co, ci, 3, 3, device='cuda') / (ci * 9) ** 0.5
    xs, ws = C.pack_split32(x), C.pack_weights(w)
    out = torch.empty(B, r, r, co, device='cuda')
    for _ in range(2):
        C.conv2d(xs, ws, 3, 3, padding=1, out=out)
    torch.cuda.synchronize()
    del x, xs, out
