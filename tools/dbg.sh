#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
WGS_EPI_FRAG=$v ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/dbg_frag$v python tools/ncu_step.py 2 "" wgs_conv_split32 > gpurun_out/dbg_frag$v.log 2>&1
tail -3 gpurun_out/dbg_frag$v.log
done
