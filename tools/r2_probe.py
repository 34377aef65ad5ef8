"""Round-2 probe (GPU): MN-major tcgen05 kernels against torch, case by case, without stopping at the first failure.

    python tools/r2_probe.py [wgrad] [linear] [time]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import ops  # noqa: E402


def rel(got, want):
    want = want.double()
    return ((got.double().cpu() - want.cpu()).norm() / max(want.norm().item(), 1e-30)).item()


def probe_wgrad():
    print("== wgrad_f16: out = alpha * a^T b over tokens ==")
    g = torch.Generator().manual_seed(0)
    cases = [(256, 128, 64, 0), (64, 64, 64, 0), (100, 64, 64, 0), (1000, 64, 64, 0), (777, 256, 64, 0), (3000, 192, 64, 0),
             (5000, 320, 1280, 0), (50176, 256, 64, 0), (50176, 64, 256, 0), (6272, 16, 64, 0), (12544, 512, 128, 0),
             (784, 1280, 320, 0), (784, 64, 64, 16), (3136, 128, 128, 16), (196, 320, 320, 16)]
    for T, NL, KL, batch in cases:
        for dt in (torch.float32, torch.float16):
            _probe_wgrad_case(g, T, NL, KL, batch, dt)


def _probe_wgrad_case(g, T, NL, KL, batch, dt):
    shp_a = (batch, T, NL) if batch else (T, NL)
    shp_b = (batch, T, KL) if batch else (T, KL)
    a = (torch.randn(shp_a, generator=g) * 0.5).to(dt)
    b = torch.randn(shp_b, generator=g).to(dt)
    want = torch.matmul(a.double().transpose(-1, -2), b.double()) * 0.25
    ch = NL // 8 if batch else 0
    try:
        out, outT, db = ops.wgrad_mn(a.cuda(), b.cuda(), alpha=0.25, need_db=not batch, need_T=bool(batch), mask_ch=ch)
        torch.cuda.synchronize()
        if batch:
            want = want * (torch.arange(NL)[:, None] // ch == torch.arange(KL)[None, :] // ch).double()
        msg = "rel %.2e" % rel(out, want)
        if outT is not None:
            msg += "  T-copy equal %s" % bool(torch.equal(outT.cpu(), out.cpu().transpose(-1, -2)))
        if db is not None:
            msg += "  db rel %.2e" % rel(db, a.double().sum(0) * 0.25)
        out2, _, db2 = ops.wgrad_mn(a.cuda(), b.cuda(), alpha=0.25, need_db=not batch, need_T=bool(batch), mask_ch=ch)
        msg += "  reproducible %s" % bool(torch.equal(out, out2) and (db is None or torch.equal(db, db2)))
    except Exception as e:  # noqa: BLE001
        msg = "FAILED: %s" % e
    print("  T=%6d NL=%4d KL=%4d batch=%2d %-8s: %s" % (T, NL, KL, batch, str(dt)[6:], msg), flush=True)


def probe_linear():
    print("== linear_bwd: bf16-gradient path / scaled-fp16 path / round-1 packT path vs fp64 torch ==")
    g = torch.Generator().manual_seed(1)
    modes = (("tf32 in place", {"wgrad_tc": 1}), ("round-1 packT", {"wgrad_tc": 0}))
    for M, N, K, x16, dscale in [(1000, 256, 64, True, 1e-4), (1000, 256, 64, True, 3.0), (777, 64, 256, True, 1e-4),
                                 (50, 2048, 512, False, 1e-4), (6272, 128, 512, True, 1e-4), (33, 320, 1280, False, 1.0),
                                 (50176, 64, 64, True, 1e-4), (50176, 256, 64, True, 1e-6), (12544, 192, 64, False, 1e-4),
                                 (784, 960, 320, True, 1e-4), (802816, 16, 64, False, 1e-6)]:
        x = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) * K ** -0.5
        dy = torch.randn(M, N, generator=g) * dscale
        if x16:
            x = x.half().float()
        want_dx = dy.double() @ w.double()
        want_dw = dy.double().t() @ x.double()
        want_db = dy.double().sum(0)
        xg = x.cuda().half() if x16 else x.cuda()
        for name, flags in modes:
            for k, v in flags.items():
                ops.set_flag(k, v)
            try:
                dx, dw, db = ops.linear_bwd(xg, w.cuda(), dy.cuda())
                torch.cuda.synchronize()
                msg = "dx %.2e dw %.2e db %.2e" % (rel(dx, want_dx), rel(dw, want_dw), rel(db, want_db))
            except Exception as e:  # noqa: BLE001
                msg = "FAILED: %s" % e
            print("  M=%6d N=%4d K=%4d x16=%d |dy|~%.0e %s: %s" % (M, N, K, x16, dscale, name, msg), flush=True)
    ops.set_flag("wgrad_tc", 1)


def _time(fn, n=20):
    """GPU time per call in us: n calls captured in one CUDA graph (no host launch cost), best of 5 replays"""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def probe_time():
    print("== linear_bwd GPU time per call (us, 20 calls in one CUDA graph) ==")
    for M, N, K in [(50176, 256, 64), (50176, 64, 256), (50176, 64, 64), (12544, 192, 64), (12544, 512, 128), (12544, 128, 512),
                    (3136, 384, 128), (3136, 128, 512), (784, 1280, 320), (784, 320, 1280), (97216, 64, 64)]:
        x = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda")
        dy = torch.randn(M, N, device="cuda") * 1e-4
        res = []
        for flag in (1, 0):
            ops.set_flag("wgrad_tc", flag)
            res.append(_time(lambda: ops.linear_bwd(x, w, dy)))
        ops.set_flag("wgrad_tc", 1)
        xf = x.float()
        tw = _time(lambda: ops.wgrad_mn(dy, xf, need_db=True))
        res.append(_time(lambda: ops.linear_bwd(xf, w, dy)))
        td = _time(lambda: ops.linear_bwd(x, w, dy, need_dw=False, need_db=False))
        print("  M=%6d N=%4d K=%4d : new (x fp16) %6.1f  (x fp32) %6.1f  old %6.1f   wgrad kernel alone %5.1f   dgrad alone %5.1f"
              % (M, N, K, res[0], res[2], res[1], tw, td), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["wgrad", "linear", "time"]
    t0 = time.time()
    print(torch.cuda.get_device_name(0))
    if "wgrad" in what:
        probe_wgrad()
    if "linear" in what:
        probe_linear()
    if "time" in what:
        probe_time()
    print("done in %.1f s" % (time.time() - t0))
