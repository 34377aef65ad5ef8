#!/usr/bin/env python
"""One eager bs16 forward for ncu: warm up outside the profiled range, then run `--iters` forwards between
cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Not a benchmark: numbers under ncu are never reported."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--flag", action="append", default=[], help="name=0/1 library flag")
    a = ap.parse_args()
    from transception_b200 import MSTransception, ops
    ops.load_library()
    for f in a.flag:
        k, v = f.split("=")
        ops.set_flag(k, int(v))
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).eval().cuda()
    x = (torch.rand(a.batch, 1, 224, 224, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
    with torch.no_grad():
        for _ in range(a.warmup):
            net(x)
        torch.cuda.synchronize()
        n0 = ops.launches()
        torch.cuda.cudart().cudaProfilerStart()
        for _ in range(a.iters):
            y = net(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("kernels per forward:", (ops.launches() - n0) // a.iters, "logits", tuple(y.shape), float(y.abs().max()))


if __name__ == "__main__":
    main()
