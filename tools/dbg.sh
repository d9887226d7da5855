#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_attribute_space.py -q -m gpu 2>&1 | tail -15
python - <<'PY'
import torch, time
from warpedganspace_b200.eval_resnet import fairface_resnet34, hopenet_resnet50
for name, mk in (('fairface resnet34', fairface_resnet34), ('hopenet resnet50', hopenet_resnet50)):
    net = mk().cuda()
    x = torch.randn(33, 3, 224, 224, device='cuda')
    for _ in range(3): net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net(x)
    e1.record(); torch.cuda.synchronize()
    print('%s: %.2f ms per batch of 33 crops (eager)' % (name, e0.elapsed_time(e1) / 10))
PY
