#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration csv): per-kernel totals and shares; optional per-launch dump."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hi]
    ki, vi, mi, ui, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Grid Size')
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        name = re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '').replace('void ', '')
        out.append((name, r[gi], v))
    return out


def main():
    L = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, v in L:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v for _, _, v in L)
    print('total %.1f us over %d launches' % (tot, len(L)))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-44s n=%4d %9.1f us %5.1f%% avg %7.1f' % (k[:44], n, t, 100 * t / tot, t / n))
    if len(sys.argv) > 2:
        for i, (n, g, v) in enumerate(L):
            if sys.argv[2] in n or sys.argv[2] == 'all':
                print('%4d %-36s %-16s %8.1f' % (i, n[:36], g, v))


if __name__ == '__main__':
    main()
