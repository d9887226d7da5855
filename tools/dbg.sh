#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_probe tools/probes/tmem_ld_layout.cu 2>/dev/null && timeout 60 /tmp/tmem_probe > gpurun_out/tmem_probe.txt 2>&1; head -40 gpurun_out/tmem_probe.txt
python -m pytest tests -q -m gpu > gpurun_out/dbg_tests.log 2>&1; tail -4 gpurun_out/dbg_tests.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
for c in c2 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-230; done
