"""profiles/rNN_bench_lines.jsonl -> a readable summary (headline line, roofline, per-layer table, the other configs).
Usage: python tools/bench_summary.py profiles/r02_bench_lines.jsonl > profiles/r02_bench_n1.md"""
import json
import sys

lines = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')]
ours = [l for l in lines if 'impl' not in l]
ref = {l['impl'] + ':' + l['metric']: l for l in lines if 'impl' in l}
d = ours[0]
r = d['roofline']
out = ['# Headline bench line (config c3, one B200)\n']
out.append('`python bench.py --steps %d --warmup %d`: **%.1f image-pairs/s** device-resident (%.2f ms per step, CUDA graph), '
           '**%.1f pairs/s end to end** (pinned-host latents in, loss out: %d B in, %d B out per step); SM clock %s / %s MHz under load, '
           'reasons %s; %d library launches per step.\n'
           % (d['steps'], d['warmup'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'],
              d['e2e']['d2h_bytes_per_step'], d['clocks']['sm_mhz'], d['clocks']['sm_max_mhz'], d['clocks']['reasons'], d['gpu_launches']))
out.append('Roofline of the tensor-core conv family: %.1f TFLOP/s algorithmic = %.3f of the sustained bf16 peak (%.1f); %s; '
           'share of the step %.2f; DRAM traffic per launch %.0f MB measured vs %.0f MB algorithmic; whole step %.1f TFLOP/s algorithmic.\n'
           % (r['achieved'], r['frac'], r['peak'], r['precision'], r['share_of_step'], (r['traffic'] or 0) / 1e6,
              r['algorithmic_bytes_per_launch'] / 1e6, r['whole_step_algorithmic_tflops']))
for key, l in ref.items():
    if l['metric'] == d['metric']:
        out.append('* `--impl %s`: %.2f %s (%.1f ms per step)' % (l['impl'], l['value'], l['unit'], l.get('ms_per_step', 0)))
out.append('\n| layer (launch, timed alone on one stream) | launches / step | us / launch | algorithmic TFLOP/s | contracted TFLOP/s | ms / step |\n|---|---|---|---|---|---|')
for row in r['per_layer']:
    out.append('| %s | %g | %.1f | %.1f | %.1f | %.3f |' % (row['layer'], row['launches_per_step'], row['us_per_launch'],
                                                          row['algorithmic_tflops'], row['contracted_tflops'], row['ms_per_step']))
out.append('\n## Other configs\n')
for l in ours[1:]:
    g = ref.get('reference-gpu:' + l['metric'])
    out.append('* %s: **%.1f %s** (%.2f ms per step), end to end %.1f, conv roofline %.3f%s'
               % (l['metric'], l['value'], l['unit'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'],
                  '; reference algorithm on the same GPU (cuDNN, TF32): %.1f' % g['value'] if g else ''))
print('\n'.join(out))
