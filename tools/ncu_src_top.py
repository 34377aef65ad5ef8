#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an .ncu-rep source page (`ncu -i … --page source --csv`)."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
si, ni = h.index('Source'), h.index('# Samples')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
body = [r for r in rows[hi + 1:] if len(r) > ni and r[0].startswith('0x')]
tot = sum(int(r[ni] or 0) for r in body)
print('total samples', tot, 'instructions', len(body))
agg = collections.Counter()
for r in body:
    for i in stall_cols:
        agg[h[i]] += int(r[i] or 0)
print('by reason:', ', '.join('%s=%d' % kv for kv in agg.most_common(10)))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ni] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    why = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
    print('%5d %6s  %-70s %s' % (i, r[ni], r[si].strip()[:70], ' '.join('%s=%d' % (n, v) for v, n in why if v)))
