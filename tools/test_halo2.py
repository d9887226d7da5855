import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys; sys.path.insert(0, %r)
import torch, torch.nn.functional as F
from warpedganspace_b200 import conv as C
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
g = torch.Generator().manual_seed(0)
for (N, Ci, H, W, Co, k, pad) in [(1, 32, 32, 32, 32, 5, 2), (1, 32, 16, 8, 32, 5, 2), (1, 32, 64, 64, 32, 5, 2), (1, 32, 32, 32, 32, 7, 3), (1, 32, 32, 32, 32, 4, 1), (1, 64, 32, 32, 32, 5, 2)]:
    x = torch.randn(N, Ci, H, W, generator=g).cuda(); w = torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** .5
    want = F.conv2d(x, w, padding=pad).permute(0, 2, 3, 1)
    got = C.conv2d(C.pack_split32(x.permute(0, 2, 3, 1).contiguous()), C.pack_weights(w), k, k, padding=pad)
    torch.cuda.synchronize()
    err = (got - want).abs()
    bad = (err > 1e-3).nonzero()
    print('  case', (N, Ci, H, W, Co, k), 'rel err %%.2e' %% rel(got, want), 'n_bad', bad.shape[0], bad[:6].tolist())
''' % ROOT
for halo in ('0', '1'):
    env = dict(os.environ, WGS_CONV_HALO=halo)
    print('halo=%s' % halo, flush=True)
    out = subprocess.run([sys.executable, '-c', CODE], env=env, capture_output=True, text=True, timeout=120)
    print(out.stdout[-2500:], out.stderr[-600:] if out.returncode else '', flush=True)
