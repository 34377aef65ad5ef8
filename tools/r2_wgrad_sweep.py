"""wgrad_tc time vs. the CTA cap (flag max_ctas limits the split plan): is a launch bound by streaming (time ~ 1 / CTAs) or by
fixed costs (flat)?   python tools/r2_wgrad_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import ops  # noqa: E402


def _time(fn, n=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


for (M, N, K) in ((50176, 256, 64), (50176, 64, 256), (97216, 64, 64), (12544, 512, 128), (3136, 384, 128), (3136, 128, 512), (784, 1280, 320)):
    dy, x = torch.randn(M, N, device="cuda") * 1e-3, torch.randn(M, K, device="cuda")
    row = []
    for cap in (0, 128, 96, 64, 32, 16, 8):
        ops.set_flag("max_ctas", cap)
        for db in (True, False):
            row.append("%s%s %.1f" % (cap or "auto", "+db" if db else "", _time(lambda: ops.wgrad_mn(dy, x, need_db=db))))
    ops.set_flag("max_ctas", 0)
    print("M=%6d N=%4d K=%4d (%.0f MB): %s" % (M, N, K, M * (N + K) * 4 / 1e6, "  ".join(row)), flush=True)
