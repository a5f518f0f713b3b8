#!/usr/bin/env python
"""Summarise an ncu --csv launch list: per launch name, grid, duration and DRAM bytes.
usage: python profiles/launch_summary.py gpurun_out/launches.csv"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
d = OrderedDict()
for r in rows[1:]:
    r = dict(zip(hdr, r))
    k = (int(r['ID']), r['Kernel Name'].replace('bk::', '')[:44], r['Grid Size'], r['Block Size'])
    d.setdefault(k, {})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
tot = sum(v['gpu__time_duration.sum'] for v in d.values())
print(f"{'id':>3} {'kernel':44s} {'grid':>14s} {'us':>9s} {'share':>6s} {'rdMB':>8s} {'wrMB':>8s}")
for k, v in d.items():
    t = v['gpu__time_duration.sum']
    print(f"{k[0]:3d} {k[1]:44s} {k[2]:>14s} {t/1e3:9.1f} {100*t/tot:5.1f}% {v.get('dram__bytes_read.sum',0)/1e6:8.1f} {v.get('dram__bytes_write.sum',0)/1e6:8.1f}")
print(f"total {tot/1e3:.1f} us")
