#!/bin/bash
# Quick iteration pass on ONE B200: GPU tests, the headline bench line, the ncu launch list of one eager step.
tag=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_tests.log 2>&1; tail -4 gpurun_out/${tag}_tests.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err; cut -c1-220 gpurun_out/${tag}_c3.json; tail -3 gpurun_out/${tag}_c3.err
if [ "$2" != "nolist" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 45 > gpurun_out/${tag}_launch_summary.md 2>&1; head -50 gpurun_out/${tag}_launch_summary.md
fi
