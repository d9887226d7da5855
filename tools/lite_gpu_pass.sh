#!/bin/bash
# Short closing pass on ONE B200 after a kernel change: GPU tests, the headline bench line, the launch list and the conv DRAM
# traffic capture that bench.py's roofline.traffic reads (gpurun -- 'bash tools/lite_gpu_pass.sh TAG').
tag=${1:-lite}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
ncu --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --csv --log-file gpurun_out/${tag}_conv_traffic.csv python tools/ncu_step.py 400 "" wgs_conv_split32 > gpurun_out/${tag}_traffic.log 2>&1
python tools/ncu_traffic.py gpurun_out/${tag}_conv_traffic.csv gpurun_out/${tag}_conv_traffic.json > /dev/null 2>&1; head -c 200 gpurun_out/${tag}_conv_traffic.json; echo
cp gpurun_out/${tag}_conv_traffic.json profiles/r02_conv_traffic.json
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err; cut -c1-200 gpurun_out/${tag}_c3.json
python bench.py --impl reference --steps 3 --warmup 1 --reference-seconds 30 > gpurun_out/${tag}_ref_c3.json 2>/dev/null; cut -c1-160 gpurun_out/${tag}_ref_c3.json
python bench.py --impl reference-gpu --steps 5 --warmup 2 > gpurun_out/${tag}_refgpu_c3.json 2>/dev/null; cut -c1-160 gpurun_out/${tag}_refgpu_c3.json
for c in c2 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err; cut -c1-160 gpurun_out/${tag}_$c.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 40 > gpurun_out/${tag}_launch_summary.md 2>&1; head -14 gpurun_out/${tag}_launch_summary.md
