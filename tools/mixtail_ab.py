#!/usr/bin/env python
"""A/B of the fused Mix-FFN tail (flag "mixtail") on the model's own Mix-FFN modules at the bs16 shapes, plus the
whole forward under a CUDA graph with the flag on and off (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transception_b200 import MSTransception, ops  # noqa: E402
from transception_b200.runtime import GraphRunner  # noqa: E402
from microbench import timeit  # noqa: E402


def main():
    torch.manual_seed(0)
    net = MSTransception(num_classes=9).eval().cuda()
    B = 16
    cases = [("backbone.block1.0.mlp", 56, 64), ("backbone.mhca_stage2.mhca_blks.0.MHCA_layers.0.mlp", 28, 64), ("backbone.mhca_stage3.mhca_blks.0.MHCA_layers.0.mlp", 14, 128),
             ("bridge.bridge_layer1.mixffn1", 56, 64), ("bridge.bridge_layer1.mixffn2", 28, 128)]
    with torch.no_grad():
        for path, hw, C in cases:
            try:
                mod = net.get_submodule(path)
            except AttributeError as e:
                print("skip", path, e)
                continue
            x = torch.randn(B, hw * hw, C, device="cuda")
            out = {}
            for flag in (0, 1):
                ops.set_flag("mixtail", flag)
                y = mod(x, hw, hw)
                t = timeit(lambda: mod(x, hw, hw))
                out[flag] = (t, y.clone())
            d = (out[0][1] - out[1][1]).abs().max().item()
            print("%-55s hw=%2d C=%3d : unfused %7.1f us  fused %7.1f us   max|diff| %.2e" % (path, hw, C, out[0][0], out[1][0], d))
        for flag in (0, 1):
            ops.set_flag("mixtail", flag)
            r = GraphRunner(net, B, 1, 224, "cuda", warmup=3)
            xs = torch.rand(B, 1, 224, 224, device="cuda")
            r.x.copy_(xs)
            for _ in range(5):
                r.graph.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(30):
                r.graph.replay()
            e.record()
            torch.cuda.synchronize()
            print("whole forward bs16 graph, mixtail=%d : %.3f ms  (%d kernels)  logits sum %.6f" % (
                flag, s.elapsed_time(e) / 30, r.kernels_per_replay, r.y.float().sum().item()))
        ops.set_flag("mixtail", 1)


if __name__ == "__main__":
    main()
