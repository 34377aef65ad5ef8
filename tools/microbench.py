#!/usr/bin/env python
"""Warm-L2 CUDA-event timings of individual library ops at the bs16 shapes (development aid, not the bench)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transception_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    """CUDA-graph replay of `iters` back-to-back calls (no host launch overhead in the number)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def timeit_eager(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def main():
    ops.load_library()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = "cuda"
    r = lambda *s: torch.randn(*s, device=dev)
    rows = []
    for (M, N, K) in [(50176, 64, 64), (50176, 256, 64), (50176, 64, 256), (12544, 64, 64), (12544, 192, 64),
                      (12544, 512, 128), (3136, 384, 128), (3136, 128, 512), (784, 1280, 320), (784, 320, 1280),
                      (97216, 64, 64), (97216, 192, 64), (784, 512, 1280)]:
        x, w, b, res = r(M, K), r(N, K), r(N), r(M, N)
        if "linear".startswith(only) or only in "linear":
            t0 = timeit(lambda: ops.linear(x, w, b))
            t1 = timeit(lambda: ops.linear(x, w, b, residual=res))
            t2 = timeit(lambda: ops.linear(x, w, b, act=1))
            gb = (M * K + N * K + M * N) * 4 / 1e9
            rows.append("linear M=%6d N=%4d K=%4d : %7.1f us  (+res %7.1f, gelu %7.1f)  min-bytes %.1f MB -> %.0f GB/s, %.1f TF/s" % (
                M, N, K, t0, t1, t2, gb * 1e3, gb / (t0 * 1e-6), 2.0 * M * N * K / (t0 * 1e-6) / 1e12))
    for (M, N, K) in [(50176, 256, 64), (50176, 64, 256), (12544, 192, 64), (12544, 512, 128), (3136, 128, 512),
                      (784, 1280, 320), (97216, 64, 64)]:
        x, w, b, res = r(M, K).half(), r(N, K).half(), r(N), r(M, N)
        if "linear".startswith(only) or only in "linear":
            t0 = timeit(lambda: ops.linear_f16(x, w, b, out_f16=True))
            t1 = timeit(lambda: ops.linear_f16(x, w, b, residual=res))
            rows.append("linear_f16 M=%6d N=%4d K=%4d : f16 out %7.1f us   f32 out + res %7.1f us" % (M, N, K, t0, t1))
    for (M, C) in [(50176, 64), (97216, 64), (12544, 128), (3136, 320), (784, 512)]:
        x, w, b = r(M, C), r(C), r(C)
        t0 = timeit(lambda: ops.layernorm(x, w, b, 1e-5))
        rows.append("layernorm M=%6d C=%4d : %7.1f us -> %.0f GB/s" % (M, C, t0, 2 * M * C * 4 / 1e9 / (t0 * 1e-6)))
    for (B, hw, C) in [(16, 56, 64), (16, 28, 128), (16, 14, 320), (16, 7, 512), (48, 28, 64), (48, 14, 128), (48, 7, 320)]:
        x = r(B, hw * hw, C)
        C4 = 4 * C
        args = (r(C4, C) * C ** -0.5, r(C4), r(C4, 1, 3, 3), r(C4), r(C4), r(C4), 1e-5, r(C, C4) * C4 ** -0.5, r(C))
        t0 = timeit(lambda: ops.mixffn_skip(x, hw, hw, *args, residual=x))
        fl = 2.0 * B * hw * hw * C * C4 * 2
        rows.append("mixffn_skip B=%2d hw=%2d C=%3d : %7.1f us -> %.1f TF/s" % (B, hw, C, t0, fl / (t0 * 1e-6) / 1e12))
    q, kv = r(16, 6076, 64), r(16, 784, 128)
    t0 = timeit(lambda: ops.flash_attn(q, kv, 0.125))
    rows.append("flash_attn 16x6076x784 : %7.1f us -> %.1f TF/s" % (t0, 4.0 * 16 * 6076 * 784 * 64 / (t0 * 1e-6) / 1e12))
    q16, kv16 = q.half(), kv.half()
    t1 = timeit(lambda: ops.flash_attn_f16(q16, kv16, 0.125))
    rows.append("flash_attn_f16 16x6076x784 : %7.1f us -> %.1f TF/s (incl. V^T pack)" % (t1, 4.0 * 16 * 6076 * 784 * 64 / (t1 * 1e-6) / 1e12))
    print("\n".join(rows))


if __name__ == "__main__":
    main()
