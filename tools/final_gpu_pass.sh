#!/bin/bash
# End-of-round measurement pass on ONE B200 (gpurun -- 'bash tools/final_gpu_pass.sh TAG'): tests, sanitizers, every bench config,
# the reference-GPU arm, the ncu launch list / top-kernel capture / conv DRAM traffic.  Everything lands in gpurun_out/<TAG>_*.
tag=${1:-r2final}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_c3.json 2> gpurun_out/${tag}_c3.err
for c in c2 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 --reference-seconds 30 > gpurun_out/${tag}_ref_c3.json 2>/dev/null
for c in c3 c2 c4; do python bench.py --impl reference-gpu --config $c --steps 5 --warmup 2 > gpurun_out/${tag}_refgpu_$c.json 2> gpurun_out/${tag}_refgpu_$c.err; done
cut -c1-160 gpurun_out/${tag}_c*.json gpurun_out/${tag}_refgpu_*.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 40 > gpurun_out/${tag}_launch_summary.md 2>&1; head -12 gpurun_out/${tag}_launch_summary.md
ncu --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --csv --log-file gpurun_out/${tag}_conv_traffic.csv python tools/ncu_step.py 400 "" wgs_conv_split32 > gpurun_out/${tag}_traffic.log 2>&1
python tools/ncu_traffic.py gpurun_out/${tag}_conv_traffic.csv gpurun_out/${tag}_conv_traffic.json > /dev/null 2>&1; head -c 300 gpurun_out/${tag}_conv_traffic.json
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/${tag}_top python tools/ncu_step.py 16 > gpurun_out/${tag}_top_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/${tag}_top.ncu-rep > gpurun_out/${tag}_top_table.md 2>&1
SANITIZE_TIMEOUT=1200 bash tools/sanitize.sh memcheck tests
SANITIZE_TIMEOUT=900 bash tools/sanitize.sh racecheck tests/test_conv_gpu.py tests/test_rbf_gpu.py tests/test_stylegan2_gpu.py tests/test_reconstructor_gpu.py
