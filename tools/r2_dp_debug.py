import os, sys
import torch
import torch.nn.parallel as P
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from networks.MSTr import MSTransception

gen = torch.Generator().manual_seed(0)
x = (torch.rand(4, 1, 224, 224, generator=gen) * 2 - 1)
torch.manual_seed(1234)
a = MSTransception(num_classes=9).train().cuda(0)
reps = P.replicate(a, [0, 1])
r1 = reps[1]

def first(o):
    while isinstance(o, (list, tuple)):
        o = o[0]
    return o if torch.is_tensor(o) else None

def record(model, store):
    hs = []
    for name, m in model.named_modules():
        def hook(mod, inp, out, name=name):
            t = first(out)
            if t is not None and name not in store:
                store[name] = t.detach().float().cpu().contiguous()
        hs.append(m.register_forward_hook(hook))
    return hs

s0, s1 = {}, {}
h = record(a, s0)
ya = a(x[2:].cuda(0))
for q in h: q.remove()
h = record(r1, s1)
with torch.cuda.device(1):
    yb = r1(x[2:].cuda(1))
for q in h: q.remove()
print("logits diff %.3e" % (ya.detach().cpu() - yb.detach().cpu()).abs().max().item())
n = 0
for name in s0:
    if name in s1 and s0[name].shape == s1[name].shape:
        dmax = (s0[name] - s1[name]).abs().max().item()
        if dmax > 0:
            print("first differing module outputs:", name, "%.3e" % dmax, type(dict(a.named_modules())[name]).__name__)
            n += 1
            if n >= 6:
                break
