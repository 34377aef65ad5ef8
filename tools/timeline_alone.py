"""Which kernels of a timeline (tools/r2_timeline.py csv) run ALONE on the GPU (nothing else in flight): the dependent chain that
bounds the step.  python tools/timeline_alone.py gpurun_out/r2e_timeline.csv"""
import csv
import sys
from collections import defaultdict

rows = [(float(r[0]), float(r[0]) + float(r[1]), r[2]) for r in list(csv.reader(open(sys.argv[1])))[1:]]
rows.sort()
pts = []
for i, (s, e, n) in enumerate(rows):
    pts.append((s, 1, i))
    pts.append((e, -1, i))
pts.sort(key=lambda p: (p[0], p[1]))
active = set()
alone = defaultdict(float)
tot = defaultdict(float)
cnt = defaultdict(int)
last = pts[0][0]
for t, d, i in pts:
    if len(active) == 1:
        alone[rows[next(iter(active))][2]] += t - last
    if d == 1:
        active.add(i)
    else:
        active.discard(i)
    last = t
for s, e, n in rows:
    tot[n] += e - s
    cnt[n] += 1
print("alone total %.2f ms" % (sum(alone.values()) / 1e3))
print("%-60s %5s %10s %10s" % ("kernel", "n", "alone us", "total us"))
for n, v in sorted(alone.items(), key=lambda kv: -kv[1])[:40]:
    print("%-60s %5d %10.1f %10.1f" % (n, cnt[n], v, tot[n]))
# coarse phases: 1 ms buckets of (alone, sum-of-durations)
if len(sys.argv) > 2:
    span = rows[-1][1]
    nb = int(span // 1000) + 1
    b_alone = [0.0] * nb
    active = set(); last = pts[0][0]
    for t, d, i in pts:
        if len(active) == 1:
            b_alone[min(int(last // 1000), nb - 1)] += t - last
        if d == 1: active.add(i)
        else: active.discard(i)
        last = t
    for b in range(nb):
        names = defaultdict(float)
        for s, e, n in rows:
            if b * 1000 <= s < (b + 1) * 1000:
                names[n] += e - s
        top = sorted(names.items(), key=lambda kv: -kv[1])[:3]
        print("ms %2d alone %4.0f us | " % (b, b_alone[b]) + ", ".join("%s %.0f" % (n[:28], v) for n, v in top))
