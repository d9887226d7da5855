#!/bin/bash
for i in 1 2; do for v in 100 80 60; do echo "BN_FILL=$v"; WGS_BN_FILL=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200; done; done
