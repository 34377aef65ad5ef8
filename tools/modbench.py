#!/usr/bin/env python
"""Graph-replay timings of the model's sub-modules at bs16 (development aid): where the forward's time really goes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transception_b200 import MSTransception, ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    ops.load_library()
    for f in sys.argv[1:]:
        k, v = f.split("=")
        ops.set_flag(k, int(v))
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).eval().cuda()
    B = 16
    x = (torch.rand(B, 1, 224, 224, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda()
    bb = net.backbone
    rows = []
    with torch.no_grad():
        t, H, W = bb.patch_embed1(x)
        rows.append(("patch_embed", timeit(lambda: bb.patch_embed1(x))))
        for i, blk in enumerate(bb.block1):
            rows.append(("stage1 eff block %d" % i, timeit(lambda: blk(t, H, W))))
            t = blk(t, H, W)
        rows.append(("stage1 norm", timeit(lambda: ops.layernorm(t, bb.norm1.weight, bb.norm1.bias, bb.norm1.eps))))
        t = ops.layernorm(t, bb.norm1.weight, bb.norm1.bias, bb.norm1.eps)
        cur = t.view(B, H, W, -1)
        outs = [cur]
        for s in (2, 3, 4):
            pe = getattr(bb, 'patch_embed_stage%d' % s)
            st = getattr(bb, 'mhca_stage%d' % s)
            rows.append(("RIPM s%d" % s, timeit(lambda: pe.nhwc(cur))))
            stacked = pe.nhwc(cur)
            P, _, h, w, C = stacked.shape
            rows.append(("  resblock s%d" % s, timeit(lambda: st.InvRes.nhwc(stacked[0]))))
            branches = [list(e.MHCA_layers) for e in st.mhca_blks]
            rows.append(("  mhca_blocks s%d (L=%d)" % (s, len(branches[0])),
                         timeit(lambda: ops.mhca_blocks(stacked.view(P, B, h * w, C), h, w, branches))))
            rows.append(("MHCA stage s%d total" % s, timeit(lambda: st.nhwc(stacked))))
            cur = st.nhwc(stacked)
            outs.append(cur)
        rows.append(("backbone total", timeit(lambda: bb.nhwc(x))))
        tok = ops.bridge_regroup(outs)
        rows.append(("bridge regroup", timeit(lambda: ops.bridge_regroup(outs))))
        for i in range(4):
            lay = getattr(net.bridge, 'bridge_layer%d' % (i + 1))
            rows.append(("bridge layer %d" % (i + 1), timeit(lambda: lay(tok))))
            tok = lay(tok)
        maps = [m.permute(0, 2, 3, 1) for m in net.bridge(ops.bridge_regroup(outs))]
        b, _, _, c = maps[3].shape
        rows.append(("decoder_3", timeit(lambda: net.decoder_3(maps[3].reshape(b, -1, c)))))
        t3 = net.decoder_3(maps[3].reshape(b, -1, c))
        rows.append(("decoder_2", timeit(lambda: net.decoder_2(t3, maps[2]))))
        t2 = net.decoder_2(t3, maps[2])
        rows.append(("decoder_1", timeit(lambda: net.decoder_1(t2, maps[1]))))
        t1 = net.decoder_1(t2, maps[1])
        rows.append(("decoder_0", timeit(lambda: net.decoder_0(t1, maps[0]))))
        rows.append(("whole forward", timeit(lambda: net(x), iters=5)))
    for k, v in rows:
        print("%-28s %9.1f us" % (k, v))


if __name__ == "__main__":
    main()
