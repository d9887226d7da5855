#!/bin/bash
# Closing pass on ONE B200 when csrc/conv.cu is unchanged since the last traffic / top-kernel capture: GPU tests (twice: the
# second run is the flakiness check), every bench config with the reference arms, the launch list of one eager step.
tag=${1:-close}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_tests2.log 2>&1; tail -1 gpurun_out/${tag}_tests2.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err; cut -c1-200 gpurun_out/${tag}_c3.json
python bench.py --impl reference --steps 3 --warmup 1 --reference-seconds 30 > gpurun_out/${tag}_ref_c3.json 2>/dev/null; cut -c1-160 gpurun_out/${tag}_ref_c3.json
for c in c3 c2 c4; do python bench.py --impl reference-gpu --config $c --steps 5 --warmup 2 > gpurun_out/${tag}_refgpu_$c.json 2>/dev/null; cut -c1-160 gpurun_out/${tag}_refgpu_$c.json; done
for c in c2 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err; cut -c1-160 gpurun_out/${tag}_$c.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 40 > gpurun_out/${tag}_launch_summary.md 2>&1; head -14 gpurun_out/${tag}_launch_summary.md
