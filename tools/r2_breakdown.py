"""Where the bs16 train step goes: forward+backward of each part of the model replayed as its own CUDA graph (inputs detached,
a fixed random cotangent), plus the train-mode forward alone.  python tools/r2_breakdown.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transception_b200 import MSTransception, ops  # noqa: E402


def graph_time(fn, n=10):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    k0 = ops.launches()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    k = ops.launches() - k0
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best, k


def fb(module_fn, inputs, params):
    """forward + backward closure over detached inputs with fixed cotangents"""
    ins = [t.detach().clone().requires_grad_() for t in inputs]
    with torch.no_grad():
        outs = module_fn(*ins)
    outs = outs if isinstance(outs, (list, tuple)) else [outs]
    cots = [torch.randn_like(o) * 1e-3 for o in outs]

    def run():
        for p in params:
            p.grad = None
        for t in ins:
            t.grad = None
        o = module_fn(*ins)
        o = o if isinstance(o, (list, tuple)) else [o]
        torch.autograd.backward(list(o), cots)
    return run


def fwd_only(module_fn, inputs):
    ins = [t.detach() for t in inputs]

    def run():
        with torch.no_grad():
            module_fn(*ins)
    return run


def main():
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).cuda().train()
    bb = net.backbone
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(16, 1, 224, 224, generator=g) * 2 - 1).cuda()
    rows = []

    def add(name, fn, ins, params):
        t, k = graph_time(fb(fn, ins, list(params)))
        rows.append((name, t, k))
        print("%-42s fwd+bwd %7.3f ms  %5d library kernels" % (name, t, k), flush=True)

    # stem + stage 1
    def stage1(xx):
        t, H, W = bb.patch_embed1(xx)
        for blk in bb.block1:
            t = blk(t, H, W)
        return t
    add("patch embed + 2 efficient blocks (56x56)", stage1, [x], list(bb.patch_embed1.parameters()) + list(bb.block1.parameters()))
    with torch.no_grad():
        maps = bb.nhwc(x)
    cur = maps[0]
    for s in (2, 3, 4):
        pe, st = getattr(bb, "patch_embed_stage%d" % s), getattr(bb, "mhca_stage%d" % s)
        add("RIPM stage %d" % s, lambda c, pe=pe: pe.nhwc(c), [cur], pe.parameters())
        with torch.no_grad():
            stacked = pe.nhwc(cur)
        add("MHCA stage %d (3 branches + ResBlock + IFF)" % s, lambda z, st=st: st.nhwc(z), [stacked], st.parameters())
        enc = st.mhca_blks[0]
        B, H, W, C = stacked.shape[1:]
        add("   one MHCAEncoder branch of stage %d" % s, lambda z, enc=enc, H=H, W=W: enc(z, (H, W)), [stacked[0].reshape(B, H * W, C)], enc.parameters())
        cur = maps[s - 1]
    add("backbone (all of the above)", lambda xx: bb.nhwc(xx), [x], bb.parameters())
    tokens = torch.cat([m.reshape(m.shape[0], -1, 64) for m in maps], dim=1)
    add("bridge (4 layers)", lambda t: net.bridge(t), [tokens], net.bridge.parameters())
    lay = net.bridge.bridge_layer2
    add("   one bridge SR-attention layer", lambda t: lay(t), [tokens], lay.parameters())
    lay1 = net.bridge.bridge_layer1
    add("   the bridge channel-attention layer", lambda t: lay1(t), [tokens], lay1.parameters())
    with torch.no_grad():
        bmaps = [m.permute(0, 2, 3, 1).contiguous() for m in net.bridge(tokens)]

    def dec(m0, m1, m2, m3):
        b, _, _, c = m3.shape
        t3 = net.decoder_3(m3.reshape(b, -1, c))
        t2 = net.decoder_2(t3, m2)
        t1 = net.decoder_1(t2, m1)
        return net.decoder_0(t1, m0)
    dparams = [p for d in (net.decoder_0, net.decoder_1, net.decoder_2, net.decoder_3) for p in d.parameters()]
    add("decoder (4 layers)", dec, bmaps, dparams)
    add("whole model forward + backward", lambda xx: net(xx), [x], net.parameters())
    t, k = graph_time(fwd_only(lambda xx: net(xx), [x]))
    print("%-42s forward %7.3f ms  %5d library kernels (train mode, no autograd)" % ("whole model", t, k))


if __name__ == "__main__":
    main()
