#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_stylegan2_gpu.py tests/test_fullsize_gpu.py tests/test_step_gpu.py -q -m gpu 2>&1 | tail -3
python - <<'PY'
import torch
from warpedganspace_b200.stylegan2 import Generator
G = Generator(1024, 512, 8).cuda().eval()
z = torch.randn(8, 512, device='cuda')
for _ in range(3): w = G.get_latent(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): w = G.get_latent(z)
e1.record(); torch.cuda.synchronize()
print('mapping network forward, B = 8: %.1f us per call (eager, incl. pixelnorm launch)' % (e0.elapsed_time(e1) * 20))
PY
for i in 1 2; do python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200; done
