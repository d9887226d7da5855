#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n8.json 2> gpurun_out/n8.err
echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','cuda_graph','e2e')})
PY
tail -3 gpurun_out/n8.err | cut -c1-300
