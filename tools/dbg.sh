#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/dbg_tests.log 2>&1; tail -5 gpurun_out/dbg_tests.log
for i in 1 2; do python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200; done
