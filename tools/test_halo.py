"""Checks the halo conv variants (descriptor addressing experiments) against fp32 torch convs."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys; sys.path.insert(0, %r)
import torch, torch.nn.functional as F
from warpedganspace_b200 import conv as C
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
g = torch.Generator().manual_seed(0)
for (N, Ci, H, W, Co, k, pad) in [(1, 32, 16, 8, 32, 3, 1), (2, 32, 32, 32, 32, 3, 1), (1, 64, 64, 64, 64, 3, 1), (2, 128, 32, 32, 128, 3, 1),
                                  (1, 32, 40, 24, 48, 3, 1), (1, 64, 33, 17, 32, 3, 1), (1, 32, 32, 32, 32, 5, 2), (1, 256, 32, 32, 256, 3, 1)]:
    x = torch.randn(N, Ci, H, W, generator=g).cuda(); w = torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** .5
    want = F.conv2d(x, w, padding=pad).permute(0, 2, 3, 1)
    got = C.conv2d(C.pack_split32(x.permute(0, 2, 3, 1).contiguous()), C.pack_weights(w), k, k, padding=pad)
    torch.cuda.synchronize()
    print('  case', (N, Ci, H, W, Co, k), 'rel err %%.2e' %% rel(got, want))
''' % ROOT
for pitch16 in ('0', '1'):
    for bo in ('0', '1'):
        env = dict(os.environ, WGS_CONV_HALO='1', WGS_CONV_HALO_PITCH16=pitch16, WGS_CONV_HALO_BASE_OFFSET=bo)
        print('variant pitch16=%s base_offset=%s' % (pitch16, bo), flush=True)
        try:
            out = subprocess.run([sys.executable, '-c', CODE], env=env, capture_output=True, text=True, timeout=120)
            print(out.stdout[-1500:], out.stderr[-600:] if out.returncode else '', flush=True)
        except subprocess.TimeoutExpired:
            print('  TIMEOUT (hang)', flush=True)
