#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): name, grid, block, launches, mean/min us."""
import csv
import sys
from collections import OrderedDict

agg = OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 10 and r[0].isdigit():
        key = (r[4][:72], r[8], r[7])
        agg.setdefault(key, []).append(float(r[-1].replace(',', '')))
only = sys.argv[2] if len(sys.argv) > 2 else None
for k, v in agg.items():
    if only and only not in k[0]:
        continue
    print(f'{k[0]:74s} {k[1]:>16s} {k[2]:>13s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:9.1f} us  min={min(v) / 1e3:9.1f}')
