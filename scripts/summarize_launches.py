#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over ONE training step
(the launches between two consecutive adam_ema_kernel launches)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines = [l for l in open(path) if not l.startswith('==')]
rows = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('pidm::', '')
    rows.append((name, v))
adam = [i for i, (n, _) in enumerate(rows) if n.startswith('adam_ema_kernel')]
if len(adam) >= 2:
    rows = rows[adam[-2] + 1: adam[-1] + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f'one step: {len(rows)} launches, {tot / 1000:.3f} ms of kernel time (ncu, serialised, cold caches)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg={v[1] / v[0]:8.2f}  {k[:100]}')
