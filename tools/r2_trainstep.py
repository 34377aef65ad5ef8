"""Whole bs16 train step as one CUDA graph under different library flags (A/B):  python tools/r2_trainstep.py name=value[,name=value] ...
Each argument is one configuration, e.g.  max_ctas=0  max_ctas=64  "max_ctas=48,branch=0".  `branch` toggles the module-level
branch streams (mstr.TRAIN_BRANCH_STREAMS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception, mstr, ops  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402
from transception_b200.optim import FusedSGD  # noqa: E402
from transception_b200.runtime import TrainStepGraph  # noqa: E402


def run(cfg):
    flags = dict(kv.split("=") for kv in cfg.split(",") if kv)
    mstr.TRAIN_BRANCH_STREAMS = bool(int(flags.pop("branch", "1")))
    for k, v in flags.items():
        ops.set_flag(k, int(v))
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).cuda().train()
    opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).cuda()
    y = torch.randint(0, 9, (16, 224, 224), generator=g).cuda()
    r = TrainStepGraph(net, CeDiceLoss(9), opt, batch=16, sample=(x, y))
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    print("%-40s %7.2f ms/step  loss %.5f  library kernels/step %d" % (cfg, best, r.loss.item(), r.kernels_per_step), flush=True)
    for k in flags:
        ops.set_flag(k, {"pdl": 1, "fork": 1, "wgrad_tc": 1}.get(k, 0))
    del r, net, opt
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for cfg in sys.argv[1:] or ["max_ctas=0"]:
        run(cfg)
