"""Whole train-step graph as ONE ncu workload (SM-busy fraction, DRAM bytes of the step):
   ncu --graph-profiling graph --profile-from-start off --clock-control none --metrics <m> --csv --log-file out.csv python tools/r2_graph_ncu.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception, ops  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402
from transception_b200.optim import FusedSGD  # noqa: E402
from transception_b200.runtime import TrainStepGraph  # noqa: E402

for kv in sys.argv[1:]:
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
torch.cuda.cudart().cudaProfilerStart()
r.replay()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss %.5f" % r.loss.item())
