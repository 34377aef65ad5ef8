"""Kernel timeline of ONE replay of the captured train step (torch.profiler / CUPTI activity records: in-situ start, duration and
stream of every kernel of the graph).  Writes gpurun_out/<tag>_timeline.csv and prints per-kernel totals + the busy/idle split.
   python tools/r2_timeline.py [tag] [flag=value ...]"""
import csv
import os
import re
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception, ops  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402
from transception_b200.optim import FusedSGD  # noqa: E402
from transception_b200.runtime import TrainStepGraph  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ops.set_flag(k, int(v))
torch.manual_seed(1234)
net = MSTransception(num_classes=9).cuda().train()
opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
g = torch.Generator().manual_seed(0)
x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).cuda()
y = torch.randint(0, 9, (16, 224, 224), generator=g).cuda()
r = TrainStepGraph(net, CeDiceLoss(9), opt, batch=16, sample=(x, y))
for _ in range(3):
    r.replay()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() >= 0]
rows = []
for e in ev:
    name = e.name
    if name.startswith("Memcpy") or name.startswith("Memset"):
        short = name.split(" ")[0]
    else:
        short = re.sub(r"^void ", "", name)
        short = short.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        short = re.sub(r"\(.*", "", short)
        if "CUDAFunctor_add<float>" in name:
            short = "ATEN_ADD"
        elif "direct_copy" in name:
            short = "ATEN_COPY"
        elif "FillFunctor" in name:
            short = "ATEN_FILL"
        short = short[:60]
    rows.append((e.time_range.start, e.time_range.end, short))
rows.sort()
t0 = rows[0][0]
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/%s_timeline.csv" % tag, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["start_us", "dur_us", "kernel"])
    for s, e, n in rows:
        w.writerow(["%.2f" % (s - t0), "%.2f" % (e - s), n])
# busy / idle / concurrency
events = []
for s, e, n in rows:
    events.append((s, 1))
    events.append((e, -1))
events.sort()
cur, last, busy, conc = 0, events[0][0], 0.0, defaultdict(float)
for t, d in events:
    if cur > 0:
        busy += t - last
    conc[min(cur, 4)] += t - last
    cur += d
    last = t
span = rows[-1][1] - t0
print("kernels %d  span %.2f ms  >=1 kernel running %.2f ms (%.1f%%)  sum of durations %.2f ms" %
      (len(rows), span / 1e3, busy / 1e3, 100 * busy / span, sum(e - s for s, e, _ in rows) / 1e3))
print("time with k kernels in flight (k=4 means >=4): " + "  ".join("%d: %.2f ms" % (k, v / 1e3) for k, v in sorted(conc.items())))
tot, cnt = defaultdict(float), defaultdict(int)
for s, e, n in rows:
    tot[n] += e - s
    cnt[n] += 1
for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
    print("%-62s n=%4d %9.1f us  avg %6.1f" % (n, cnt[n], v, v / cnt[n]))
