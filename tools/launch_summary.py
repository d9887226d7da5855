"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: finds one training step (the launches
between two pairs of adam_kernel launches) and prints per-kernel totals / shares, optionally the launch sequence."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((row['Kernel Name'], float(row['Metric Value'].replace(',', '')) / 1e6, row['Grid Size']))
    return rows


def one_step(rows):
    adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r[0]]
    assert len(adam) >= 4, 'need two optimiser boundaries in the capture'
    return rows[adam[1] + 1: adam[3] + 1]


def short(name):
    return re.sub(r'\(.*', '', name).replace('void ', '')[:72]


if __name__ == '__main__':
    step = one_step(load(sys.argv[1]))
    total = sum(v for _, v, _ in step)
    print('one step = %d launches, %.3f ms summed\n' % (len(step), total))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _ in step:
        agg[short(n)][0] += 1
        agg[short(n)][1] += v
    print('| kernel | launches | total ms | share |\n|---|---|---|---|')
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
        print('| `%s` | %d | %.3f | %.1f %% |' % (n, c, v, 100 * v / total))
    if len(sys.argv) > 3:
        thr = float(sys.argv[3])
        for i, (n, v, g) in enumerate(step):
            if v > thr:
                print(i, short(n)[:44], '%.3f' % v, g)
