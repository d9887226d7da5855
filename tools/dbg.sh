#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n2.json 2> gpurun_out/n2.err
cut -c1-400 gpurun_out/n2.json; tail -5 gpurun_out/n2.err
python - <<'PY'
import json
for l in open('gpurun_out/n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','cuda_graph','e2e')})
PY
