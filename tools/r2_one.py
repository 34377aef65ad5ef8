"""Run one MN-major weight-gradient launch shape a few times (for ncu): python tools/r2_one.py T NL KL [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import ops  # noqa: E402

T, NL, KL = (int(v) for v in sys.argv[1:4])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(((batch,) if batch else ()) + (T, NL), device="cuda").half()
b = torch.randn(((batch,) if batch else ()) + (T, KL), device="cuda").half()
for _ in range(4):
    ops.wgrad_mn(a, b, need_db=not batch)
torch.cuda.synchronize()
