import sys, os
sys.path.insert(0, '/root/repo')
import torch
from transception_b200 import ops
from tools.microbench import timeit, timeit_eager
ops.load_library()
x = torch.randn(784, 64, device='cuda').half(); w = torch.randn(64, 64, device='cuda').half(); b = torch.randn(64, device='cuda')
def chain():
    y = x
    for _ in range(20):
        y = ops.linear_f16(y, w, b, out_f16=True)
    return y
for pdl in (1, 0, 1, 0):
    ops.set_flag("pdl", pdl)
    print("pdl", pdl, "graph: %.2f us per GEMM" % (timeit(chain, iters=5) / 20), " eager: %.2f us per GEMM" % (timeit_eager(chain, iters=5) / 20))
xl = torch.randn(50176, 64, device='cuda').half()
def chain2():
    y = xl
    for _ in range(10):
        y = ops.linear_f16(y, w, b, out_f16=True)
    return y
for pdl in (1, 0):
    ops.set_flag("pdl", pdl)
    print("big pdl", pdl, "graph: %.2f us per GEMM" % (timeit(chain2, iters=5) / 10))
