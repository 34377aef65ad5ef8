#!/usr/bin/env python
"""Per-launch dump of an ncu launch list: index, kernel, grid, block, time (us). Optional substring filter."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[hi]
ki, vi, mi, ui, gi, bi = (h.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit', 'Grid Size', 'Block Size'))
flt = sys.argv[2] if len(sys.argv) > 2 else ''
n = 0
for r in rows[hi + 1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    name = re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '').replace('void ', '')
    if flt in name:
        print('%4d %-40s %-14s %-12s %8.1f' % (n, name[:40], r[gi].replace(' ', ''), r[bi].replace(' ', ''), v))
    n += 1
