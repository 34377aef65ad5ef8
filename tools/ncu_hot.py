"""Top sampled instructions of an ncu --set full --import-source report: python tools/ncu_hot.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ci["# Samples"]])
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print("kernel:", rows[0][1] if rows and len(rows[0]) > 1 else "?", " total samples", tot)
agg = {s: 0 for s in stalls}
for n, r in data:
    for s in stalls:
        try:
            agg[s] += int(r[ci[s]])
        except ValueError:
            pass
print("stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
data.sort(key=lambda t: -t[0])
for n, r in data[:top]:
    why = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (n, 100.0 * n / max(tot, 1), r[ci["Source"]][:70], " ".join("%s:%d" % (w, c) for c, w in why if c)))
