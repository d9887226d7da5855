#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_gpu.py tests/test_generators_gpu.py -q -m gpu 2>&1 | tail -4
for c in c2 c3; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200; done
