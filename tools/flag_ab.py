#!/usr/bin/env python
"""A/B of one library flag on the whole bs16 forward under a CUDA graph: `python tools/flag_ab.py ea_tc [model]`
(development aid).  Prints ms per forward with the flag off / on and the max-abs difference of the logits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import transception_b200  # noqa: E402
from transception_b200 import ops  # noqa: E402
from transception_b200.runtime import GraphRunner  # noqa: E402


def main():
    flag = sys.argv[1]
    model = sys.argv[2] if len(sys.argv) > 2 else "MSTransception"
    torch.manual_seed(0)
    net = getattr(transception_b200, model)(num_classes=9).eval().cuda()
    xs = torch.rand(16, 1, 224, 224, device="cuda") * 2 - 1
    outs = {}
    for rep in range(2):
        for v in (0, 1):
            ops.set_flag(flag, v)
            r = GraphRunner(net, 16, 1, 224, "cuda", warmup=3)
            r.x.copy_(xs)
            for _ in range(5):
                r.graph.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(40):
                r.graph.replay()
            e.record()
            torch.cuda.synchronize()
            outs[v] = r.y.float().clone()
            print("%s %s=%d : %.3f ms / forward (%d kernels)" % (model, flag, v, s.elapsed_time(e) / 40, r.kernels_per_replay))
    print("max|logits(on) - logits(off)| = %.3e (absmax %.3e)" % ((outs[0] - outs[1]).abs().max().item(), outs[0].abs().max().item()))


if __name__ == "__main__":
    main()
