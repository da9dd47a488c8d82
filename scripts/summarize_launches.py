#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel totals over ONE training step (the launches between two consecutive adam_ema_kernel launches)."""
import collections
import csv
import json
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines = [l for l in open(path) if not l.startswith('==')]
launch = collections.OrderedDict()          # ID -> {'name':…, 'us':…, 'rd':…, 'wr':…}


def to_us(v, unit):
    return v / 1000 if unit in ('ns', 'nsecond') else (v * 1000 if unit in ('ms', 'msecond') else v)


def to_bytes(v, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return v * scale.get(unit, 1)


for row in csv.DictReader(lines):
    m = row.get('Metric Name')
    if m not in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum'):
        continue
    v = float(row['Metric Value'].replace(',', ''))
    rec = launch.setdefault(row['ID'], {'name': re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('pidm::', ''),
                                        'us': 0.0, 'rd': 0.0, 'wr': 0.0})
    if m == 'gpu__time_duration.sum':
        rec['us'] = to_us(v, row['Metric Unit'])
    elif m == 'dram__bytes_read.sum':
        rec['rd'] = to_bytes(v, row['Metric Unit'])
    else:
        rec['wr'] = to_bytes(v, row['Metric Unit'])
rows = list(launch.values())
adam = [i for i, r in enumerate(rows) if r['name'].startswith('adam_ema_kernel')]
if len(adam) >= 2:
    rows = rows[adam[-2] + 1: adam[-1] + 1]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in rows:
    a = agg[r['name']]
    a[0] += 1
    a[1] += r['us']
    a[2] += r['rd']
    a[3] += r['wr']
tot = sum(v[1] for v in agg.values())
tot_b = sum(v[2] + v[3] for v in agg.values())
print(f'one step: {len(rows)} launches, {tot / 1000:.3f} ms of kernel time, {tot_b / 1e9:.2f} GB of DRAM traffic '
      f'(ncu, serialised, cold caches)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    gbs = (v[2] + v[3]) / (v[1] * 1e-6) / 1e9 if v[1] > 0 else 0.0
    print(f'{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg={v[1] / v[0]:8.2f}  dram {(v[2] + v[3]) / 1e6:8.1f} MB '
          f'({gbs:6.0f} GB/s)  {k[:90]}')
if len(sys.argv) > 2:      # machine-readable per-kernel DRAM traffic of one step (read by bench.py for roofline.traffic)
    json.dump({k: {'launches': v[0], 'us': v[1], 'dram_bytes': v[2] + v[3]} for k, v in agg.items()}, open(sys.argv[2], 'w'),
              indent=1)
