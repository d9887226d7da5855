import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import test_step_gpu as T
from warpedganspace_b200.trainer import PairedTrainer
import torch.nn.functional as F
ch = {4: 64, 8: 64, 16: 32, 32: 32}
K, D, B, size = 16, 4, 4, 32
_, (W, S, R) = T.build(size, ch, K, D, 60)
tr = PairedTrainer(W, S, R)
g = T.gen(1)
z = torch.randn(B, 512, generator=g).cuda(); idx = torch.randint(0, K, (B,), generator=g).cuda(); mag = (torch.rand(B, generator=g) * 0.1 + 0.1).cuda()

def try_capture(name, fn, mode='global'):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(gph, capture_error_mode=mode):
            fn()
        gph.replay(); torch.cuda.synchronize()
        print(name, mode, 'OK')
    except Exception as e:
        print(name, mode, 'FAILED', repr(e)[:300])
        torch.cuda.synchronize()

def f_rbf():
    return tr.S.warp(idx, z, mag)
def f_gen():
    with torch.no_grad():
        return W(z)
def f_pair():
    with torch.no_grad():
        return W.forward_pair(z, tr.S.warp(idx, z, mag))
def f_rec():
    with torch.no_grad():
        x = torch.randn(B, 3, size, size, device='cuda')
        return R(x, x)
def f_fwd():
    shift = tr.S.warp(idx, z, mag)
    a, b = W.forward_pair(z, shift)
    return R(a.detach(), b)
def f_fb():
    tr.flat_s.zero_grad(); tr.flat_r.zero_grad()
    logits, pred = f_fwd()
    (F.cross_entropy(logits, idx) + 0.25 * (pred - mag).abs().mean()).backward()
def f_opt():
    tr.optimizer_step()
for name, fn in [('rbf', f_rbf), ('gen', f_gen), ('pair', f_pair), ('rec', f_rec), ('fwd', f_fwd), ('opt', f_opt), ('fwd+bwd', f_fb)]:
    try_capture(name, fn)
try_capture('fwd+bwd', f_fb, 'thread_local')
try_capture('fwd+bwd', f_fb, 'relaxed')
