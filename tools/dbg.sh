#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s -k "kink_free" 2>&1 | grep -E "^seed|pinned|passed|failed|Error" | head -30
python -m pytest tests -q -m gpu > gpurun_out/dbg_tests.log 2>&1; tail -6 gpurun_out/dbg_tests.log
for v in 1 0; do WGS_SIDE_STREAMS=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200; done
