"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): one bs2 inference forward and one bs1 training step
(forward + loss + backward + fused SGD) with PDL and stream forking on — the mbarrier / TMEM / TMA / cluster kernels included.
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception, ops  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402
from transception_b200.optim import FusedSGD  # noqa: E402

ops.set_flag("pdl", 1)
ops.set_flag("fork", 1)
torch.manual_seed(1234)
net = MSTransception(num_classes=9).cuda()
g = torch.Generator().manual_seed(0)
x = (torch.rand(2, 1, 224, 224, generator=g) * 2 - 1).cuda()
labels = torch.randint(0, 9, (2, 224, 224), generator=g).cuda()
with torch.no_grad():
    y = net.eval()(x)
torch.cuda.synchronize()
print("forward ok", tuple(y.shape), float(y.abs().mean()))
if "fwd" not in sys.argv:
    net.train()
    opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    loss = CeDiceLoss(9)(net(x[:1]), labels[:1])
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    print("train step ok, loss %.6f" % loss.item())
