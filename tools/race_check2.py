#!/usr/bin/env python
"""Where does run-to-run nondeterminism enter? Compare every stage output of two eager forwards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transception_b200 import MSTransception, ops
ops.load_library()
torch.manual_seed(1234)
net = MSTransception(num_classes=9).eval().cuda()
x = (torch.rand(2, 1, 224, 224, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
def d(a, b): return (a - b).abs().max().item()
def run():
    bb = net.backbone
    out = {}
    with torch.no_grad():
        t, H, W = bb.patch_embed1(x); out['pe'] = t.clone()
        for i, blk in enumerate(bb.block1):
            t = blk(t, H, W); out['eff%d' % i] = t.clone()
        t = ops.layernorm(t, bb.norm1.weight, bb.norm1.bias, bb.norm1.eps)
        cur = t.view(2, H, W, -1); outs = [cur]
        for s in (2, 3, 4):
            stacked = getattr(bb, 'patch_embed_stage%d' % s).nhwc(cur); out['ripm%d' % s] = stacked.clone()
            st = getattr(bb, 'mhca_stage%d' % s)
            P, B, h, w, C = stacked.shape
            res = st.InvRes.nhwc(stacked[0]); out['res%d' % s] = res.clone()
            enc = ops.mhca_blocks(stacked.view(P, B, h * w, C), h, w, [list(e.MHCA_layers) for e in st.mhca_blks]); out['mhca%d' % s] = enc.clone()
            cur = st.aggregate.nhwc([res] + [enc[i].view(B, h, w, C) for i in range(P)]); out['iff%d' % s] = cur.clone()
            outs.append(cur)
        tok = ops.bridge_regroup(outs)
        for i in range(4):
            tok = getattr(net.bridge, 'bridge_layer%d' % (i + 1))(tok); out['bridge%d' % (i + 1)] = tok.clone()
    return out
a, b = run(), run()
for k in a:
    print("%-10s %.3e" % (k, d(a[k], b[k])))
