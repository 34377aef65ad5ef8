"""One GPU, one-rank process group: the data-parallel step's machinery (backward cut, bucket gathers, captured collectives on a
one-rank group) against the plain single-GPU step — separates the cost of the machinery from the cost of real communication.
   python tools/r2_overlap_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402
from transception_b200.optim import FusedSGD  # noqa: E402
from transception_b200.runtime import TrainStepGraph  # noqa: E402


def run(tag, **kw):
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).cuda().train()
    opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).cuda()
    y = torch.randint(0, 9, (16, 224, 224), generator=g).cuda()
    r = TrainStepGraph(net, CeDiceLoss(9), opt, batch=16, sample=(x, y), **kw)
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
    print("%-46s %7.2f ms/step  loss %.5f  overlap=%s single_graph=%s" % (tag, best, r.loss.item(), r.overlap, getattr(r, "single_graph", None)), flush=True)


os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1)
torch.cuda.set_device(0)
run("plain (world 1)")
os.environ["TCX_FORCE_OVERLAP"] = "1"
run("overlap machinery, one graph", single_graph=True)
run("overlap machinery, five graphs", single_graph=False)
dist.destroy_process_group()
