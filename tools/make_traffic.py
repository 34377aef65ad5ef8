#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu launch list taken with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
over one eager train step: per profiled kernel family (the names bench.py's roofline legs use) the DRAM bytes per launch
(read + write, averaged over every launch of the step) and the summed duration.
   python tools/make_traffic.py profiles/r02_train_step_launches.csv profiles/r02_traffic.json"""
import collections
import csv
import json
import re
import sys

FAMILIES = {"gemm_tc": "gemm_tc_kernel", "wgrad_tc": "wgrad_tc_kernel", "dw_bwd_fused": "dw_bwd_sweep_kernel", "ln_bwd_fused": "ln_bwd_fused",
            "dwln": "dwln_kernel", "dwconv3x3": "dwconv3x3_kernel", "flash_bwd": "flash_bwd_kernel", "flash_tc": "flash_tc_kernel",
            "final_head_bwd": "final_head_bwd_kernel", "mb_fused16": "mb_fused16_kernel"}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hi]
    idi, ki, vi, mi, ui = h.index('ID'), h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit')
    per = collections.defaultdict(dict)
    names = {}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
        per[r[idi]][r[mi]] = v
        names[r[idi]] = re.sub(r'\(.*', '', r[ki])
    out = {}
    for fam, pat in FAMILIES.items():
        ids = [i for i, n in names.items() if pat in n]
        if not ids:
            continue
        b = sum(per[i].get('dram__bytes_read.sum', 0.0) + per[i].get('dram__bytes_write.sum', 0.0) for i in ids)
        t = sum(per[i].get('gpu__time_duration.sum', 0.0) for i in ids)
        out[fam] = {"dram_bytes_per_launch": b / len(ids), "launches": len(ids), "us_per_launch_cold_serialised": t / len(ids),
                    "source": src + " (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager bs16 train step, "
                              "every launch of the family)"}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out.items():
        print("%-16s n=%4d  %8.2f MB/launch  %7.1f us/launch" % (k, v["launches"], v["dram_bytes_per_launch"] / 1e6, v["us_per_launch_cold_serialised"]))


if __name__ == "__main__":
    main()
