#!/usr/bin/env python
"""Per kernel family of ONE eager train step: launches, summed duration, DRAM bytes read / written and the DRAM rate each family
achieves, from an ncu launch list taken with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
(cold-cache, serialised launches: the rates are lower bounds of what the kernels reach back to back inside the graph).

    python tools/dram_table.py profiles/r02_train_step_launches.csv > profiles/r02_train_step_dram_table.txt
"""
import collections
import csv
import json
import os
import re
import sys


def main():
    src = sys.argv[1]
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    idi, ki, vi, mi, ui = (h.index(k) for k in ("ID", "Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}
    per, names = collections.defaultdict(dict), {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        per[r[idi]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        n = re.sub(r"\(.*", "", r[ki])
        n = re.sub(r"^void\s+", "", n).replace("<unnamed>::", "")
        n = re.sub(r"<.*", "", n) if n.startswith("at::") else n
        names[r[idi]] = n
    fam = collections.OrderedDict()
    for i, n in names.items():
        f = fam.setdefault(n, [0, 0.0, 0.0, 0.0])
        f[0] += 1
        f[1] += per[i].get("gpu__time_duration.sum", 0.0)
        f[2] += per[i].get("dram__bytes_read.sum", 0.0)
        f[3] += per[i].get("dram__bytes_write.sum", 0.0)
    peak = 6538.0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    tot = [sum(f[k] for f in fam.values()) for k in range(4)]
    print("# %s: %d launches, %.1f ms summed kernel time (serialised), DRAM %.2f GB read + %.2f GB written per step" %
          (src, tot[0], tot[1] / 1e3, tot[2] / 1e9, tot[3] / 1e9))
    print("# DRAM rate = (read + written) / summed duration; peak = %.0f GB/s (MEASURED_PEAKS.json hbm copy, else the recipe's fallback)" % peak)
    print("# ncu flushes the caches before every profiled launch (--cache-control all, the default): reads are upper bounds (every")
    print("# operand comes from DRAM, inside the graph the producer's output is L2-resident) and most write-back happens after the")
    print("# kernel's measurement window, so the write column under-counts; the whole step as ONE workload is in r02_graph_ncu.csv")
    print("# (10.5 GB read + 16.0 GB written).")
    print("%-44s %5s %9s %6s %9s %9s %8s %6s" % ("kernel", "n", "us total", "share", "MB rd/l", "MB wr/l", "GB/s", "frac"))
    for n, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        gbs = (f[2] + f[3]) / (f[1] * 1e-6) / 1e9 if f[1] else 0.0
        print("%-44s %5d %9.1f %5.1f%% %9.2f %9.2f %8.0f %6.3f" % (n[:44], f[0], f[1], 100 * f[1] / tot[1], f[2] / f[0] / 1e6, f[3] / f[0] / 1e6, gbs, gbs / peak))
    gbs = (tot[2] + tot[3]) / (tot[1] * 1e-6) / 1e9
    print("%-44s %5d %9.1f %5.1f%% %9s %9s %8.0f %6.3f" % ("ALL", tot[0], tot[1], 100.0, "", "", gbs, gbs / peak))


if __name__ == "__main__":
    main()
