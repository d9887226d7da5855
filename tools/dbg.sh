#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_gpu.py tests/test_generators_gpu.py tests/test_step_gpu.py tests/test_stylegan2_gpu.py -q -m gpu 2>&1 | tail -12
for i in 1 2; do for v in 1 0; do echo "FUSE_PIXNORM=$v"; WGS_FUSE_PIXNORM=$v python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200; done; done
