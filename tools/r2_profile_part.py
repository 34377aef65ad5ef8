"""One fwd+bwd of a part of the model between cudaProfilerStart/Stop (for an ncu launch list):
   python tools/r2_profile_part.py bridge|decoder|stage1|mhca3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception, mstr  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "bridge"
mstr.TRAIN_BRANCH_STREAMS = False
torch.manual_seed(1234)
net = MSTransception(num_classes=9).cuda().train()
g = torch.Generator().manual_seed(0)
x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).cuda()
with torch.no_grad():
    maps = net.backbone.nhwc(x)
    tokens = torch.cat([m.reshape(m.shape[0], -1, 64) for m in maps], dim=1)
    bmaps = [m.permute(0, 2, 3, 1).contiguous() for m in net.bridge(tokens)]
    stacked3 = net.backbone.patch_embed_stage3.nhwc(maps[1])


def dec(m0, m1, m2, m3):
    b, _, _, c = m3.shape
    t3 = net.decoder_3(m3.reshape(b, -1, c))
    t2 = net.decoder_2(t3, m2)
    t1 = net.decoder_1(t2, m1)
    return net.decoder_0(t1, m0)


def stage1(xx):
    t, H, W = net.backbone.patch_embed1(xx)
    for blk in net.backbone.block1:
        t = blk(t, H, W)
    return t


fn, ins = {"bridge": (lambda t: net.bridge(t), [tokens]), "decoder": (dec, bmaps), "stage1": (stage1, [x]),
           "mhca3": (lambda z: net.backbone.mhca_stage3.nhwc(z), [stacked3])}[what]
ins = [t.detach().clone().requires_grad_() for t in ins]


def run():
    o = fn(*ins)
    o = o if isinstance(o, (list, tuple)) else [o]
    torch.autograd.backward(list(o), [torch.ones_like(t) * 1e-3 for t in o])


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
