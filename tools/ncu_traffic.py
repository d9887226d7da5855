"""Turns an ncu metrics capture of the tensor-core conv launches of one training step into
profiles/rNN_conv_traffic.json (bench.py reads the newest one for `roofline.traffic`).  The file is stamped with the sha1 of
the CUDA sources it measured (bench.source_stamp): bench.py reports `traffic_stale` instead of a number when the kernels
have changed since the capture.

  ncu --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --csv --log-file gpurun_out/conv_traffic.csv python tools/ncu_step.py 400 "" wgs_conv_split32
  python tools/ncu_traffic.py gpurun_out/conv_traffic.csv profiles/r02_conv_traffic.json
"""
import collections
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rows = collections.defaultdict(dict)
names = {}
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if 'Metric Name' not in r:
        continue
    names[r['ID']] = re.sub(r'\(.*', '', r['Kernel Name'])
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit'].lower()
    scale = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6,
             'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}.get(unit, 1)
    rows[r['ID']][r['Metric Name']] = v * scale
conv = {k: v for k, v in rows.items() if 'conv_' in names[k]}
n = len(conv)
rd = sum(v.get('dram__bytes_read.sum', 0) for v in conv.values())
wr = sum(v.get('dram__bytes_write.sum', 0) for v in conv.values())
ms = sum(v.get('gpu__time_duration.sum', 0) for v in conv.values())
per_kernel = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, v in conv.items():
    e = per_kernel[names[k]]
    e[0] += 1
    e[1] += v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)
    e[2] += v.get('gpu__time_duration.sum', 0)
out = {
    'dram_bytes_per_launch': (rd + wr) / max(1, n), 'launches': n, 'dram_read_bytes': rd, 'dram_write_bytes': wr,
    'summed_ms': ms,
    'source_stamp': __import__('bench').source_stamp(),
    'per_kernel': {k: {'launches': c, 'dram_bytes': b, 'ms': t} for k, (c, b, t) in per_kernel.items()},
    'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum over every tensor-core conv launch of one eager training step '
              '(tools/ncu_step.py, B = 4 per GPU), averaged per launch',
}
with open(sys.argv[2], 'w') as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
