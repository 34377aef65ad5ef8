"""Kernel timeline of ONE data-parallel train step on rank 0 (torchrun, >= 2 ranks): where the NCCL kernels sit relative to the
backward.   torchrun --nproc-per-node 2 tools/r2_timeline_dp.py [single_graph 0|1]"""
import os
import re
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402
from transception_b200.optim import FusedSGD  # noqa: E402
from transception_b200.runtime import TrainStepGraph  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
single = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
torch.manual_seed(1234)
net = MSTransception(num_classes=9).cuda().train()
opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
g = torch.Generator().manual_seed(rank)
x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).cuda()
y = torch.randint(0, 9, (16, 224, 224), generator=g).cuda()
r = TrainStepGraph(net, CeDiceLoss(9), opt, batch=16, sample=(x, y), single_graph=single)
for _ in range(3):
    r.replay()
torch.cuda.synchronize()
dist.barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r.replay()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    rows = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
    t0 = rows[0][0]
    print("single_graph=%s overlap=%s kernels %d span %.2f ms" % (r.single_graph, r.overlap, len(rows), (rows[-1][1] - t0) / 1e3))
    for s, e, n in rows:
        if "nccl" in n.lower() or "mt_gather" in n or "mt_sgd" in n or "final_head_bwd" in n or "coord_gate_bwd" in n or "patch_im2row" in n:
            print("%9.1f %8.1f %s" % (s - t0, e - s, re.sub(r"\(.*", "", n)[:70]))
dist.barrier()
dist.destroy_process_group()
