#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/s3_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s3_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/s3_launches.csv 24 2>&1 | head -30
