"""Profiles the TOP-K most expensive C-ABI launches of one eager paired step.
Pass 1 times every `_lib.call` with CUDA events; pass 2 brackets the chosen calls with cudaProfilerStart/Stop, so
`ncu --profile-from-start off --set full -o rep python tools/ncu_step.py [K] [skip_names]` captures only those."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from warpedganspace_b200 import _lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
SKIP = set(sys.argv[2].split(',')) if len(sys.argv) > 2 else set()
dev = torch.device('cuda', 0)
tr = bench.build_product(dev, 4)
bs = bench.make_batches(4, 4, dev, 1)
for i in range(2):
    tr.step(*bs[i], eager=True)
torch.cuda.synchronize()

orig = _lib.call
log, mode, chosen, idx = [], 'time', set(), [0]

def wrapped(name, *args):
    i = idx[0]; idx[0] += 1
    if mode == 'time':
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(name, *args); e1.record(); e1.synchronize()
        log.append((e0.elapsed_time(e1), i, name))
        return r
    if i in chosen:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
        r = orig(name, *args)
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
        return r
    return orig(name, *args)

_lib.call = wrapped
# the conv entry point is called through _lib.check(lib.wgs_conv_split32(...)), not _lib.call: wrap it one level up
from warpedganspace_b200 import conv as C
_conv_taps = C.conv_taps
def conv_wrapped(*a, **k):
    global orig
    saved, orig = orig, (lambda name, *aa: _conv_taps(*a, **k))
    try:
        return wrapped('wgs_conv_split32')
    finally:
        orig = saved
C.conv_taps = conv_wrapped
ONLY = set(sys.argv[3].split(',')) if len(sys.argv) > 3 and sys.argv[3] else None
MAX_MS = float(sys.argv[4]) if len(sys.argv) > 4 else 1e9          # only launches cheaper than this (the latency-bound ones)
tr.step(*bs[2], eager=True)
torch.cuda.synchronize()
top = sorted([l for l in log if l[2] not in SKIP and (ONLY is None or l[2] in ONLY) and l[0] < MAX_MS], reverse=True)[:K]
for ms, i, name in top:
    print('%4d %-28s %.3f ms' % (i, name, ms))
print('all calls: %d, total %.2f ms' % (len(log), sum(l[0] for l in log)))
chosen = set(i for _, i, _ in top)
mode, idx[0] = 'prof', 0
tr.step(*bs[3], eager=True)
torch.cuda.synchronize()
