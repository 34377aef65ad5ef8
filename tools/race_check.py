#!/usr/bin/env python
"""Determinism / hazard check: eager vs eager, graph vs graph, graph vs eager, under the pdl / fork flags."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transception_b200 import MSTransception, ops
from transception_b200.runtime import GraphRunner
ops.load_library()
torch.manual_seed(1234)
net = MSTransception(num_classes=9).eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = (torch.rand(B, 1, 224, 224, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
def d(a, b): return (a - b).abs().max().item()
for pdl, fork in ((1, 1), (0, 1), (1, 0), (0, 0)):
    ops.set_flag("pdl", pdl); ops.set_flag("fork", fork)
    with torch.no_grad():
        e1 = net(x).clone(); e2 = net(x).clone()
    r = GraphRunner(net, B, 1, 224)
    r.x.copy_(x); r.replay(); torch.cuda.synchronize(); g1 = r.y.clone()
    r.replay(); torch.cuda.synchronize(); g2 = r.y.clone()
    print("pdl=%d fork=%d  eager-eager %.2e  graph-graph %.2e  graph-eager %.2e" % (pdl, fork, d(e1, e2), d(g1, g2), d(g1, e1)))
ops.set_flag("pdl", 0); ops.set_flag("fork", 0)
with torch.no_grad():
    base = net(x).clone()
ops.set_flag("pdl", 1); ops.set_flag("fork", 1)
with torch.no_grad():
    full = net(x).clone()
print("eager(pdl,fork) vs eager(plain): %.2e" % d(base, full))
