"""Direct GPU tests of the round-2 training kernels that the module-level suites (test_gpu_backward.py) reach only through
whole modules: the fused class head, the bridge slab split / merge, Scale_reduce as one node, the stem conv node, the dual-output
LayerNorm, the fan-out node of shared parameters.  Each against torch autograd over the reference's own op sequence
(MSTr.py lines cited per test) on seeded inputs; every backward also bit-reproducible run to run."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _close(got, want, rel, what, floor=0.0):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.isfinite(got).all(), what
    err, den = (got - want).norm().item(), want.norm().item()
    assert err <= rel * den + floor, "%s: relative L2 %.3e > %.1e" % (what, err / max(den, 1e-30), rel)


@pytest.mark.parametrize("ncls", [1, 2, 9, 16])
def test_final_head_forward_backward(cuda_lib, ncls):
    """FinalPatchExpand_X4's rearrange + LayerNorm(64) (MSTr.py:212-227) + last_layer 1x1 conv (:288-289) on the expand output."""
    from transception_b200 import autograd as A
    B, H, W = 2, 5, 7
    e = _rand(B, H * W, 1024, seed=1)
    lnw, lnb = 1 + 0.2 * _rand(64, seed=2), 0.1 * _rand(64, seed=3)
    cw, cb = 0.2 * _rand(ncls, 64, 1, 1, seed=4), 0.1 * _rand(ncls, seed=5)
    dy = _rand(B, ncls, 4 * H, 4 * W, seed=6)

    def ref(e, lnw, lnb, cw, cb):
        x = e.view(B, H, W, 4, 4, 64).permute(0, 1, 3, 2, 4, 5).reshape(B, 4 * H * 4 * W, 64)      # 'b h w (p1 p2 c) -> b (h p1) (w p2) c'
        x = F.layer_norm(x, (64,), lnw, lnb, 1e-5)
        return F.conv2d(x.view(B, 4 * H, 4 * W, 64).permute(0, 3, 1, 2), cw, cb)
    r = [t.clone().requires_grad_() for t in (e, lnw, lnb, cw, cb)]
    yr = ref(*r)
    yr.backward(dy)
    g = [t.cuda().requires_grad_() for t in (e, lnw, lnb, cw, cb)]
    y = A.final_head(g[0], H, W, g[1], g[2], 1e-5, g[3], g[4])
    _close(y, yr, 1e-5, "logits")
    y.backward(dy.cuda())
    for name, a, b in zip(("de", "d ln_w", "d ln_b", "d cls_w", "d cls_b"), g, r):
        _close(a.grad, b.grad, 2e-5, name)
    first = [t.grad.clone() for t in g]
    for t in g:
        t.grad = None
    A.final_head(g[0], H, W, g[1], g[2], 1e-5, g[3], g[4]).backward(dy.cuda())
    assert all(torch.equal(a, t.grad) for a, t in zip(first, g))


def test_bridge_split_merge_are_adjoint_copies(cuda_lib):
    """MSTr.py:2380-2386 / :2394-2403: the token buffer [B, Ntok, 64] and its four slabs; merge(split(t)) == t, the gradient of
    one is the other, and the residual of the merge receives the gradient unchanged."""
    from transception_b200 import autograd as A
    B, S = 3, 16
    n = [S * S, (S // 2) ** 2 * 2, (S // 4) ** 2 * 5, (S // 8) ** 2 * 8]
    t = _rand(B, sum(n), 64, seed=1).cuda().requires_grad_()
    slabs = A.bridge_split(t)
    off = 0
    for k, (s, nk) in enumerate(zip(slabs, n)):
        assert s.shape == (B, (S >> k) ** 2, 64 * (1, 2, 5, 8)[k])
        assert torch.equal(s.reshape(B, nk, 64), t[:, off:off + nk])
        off += nk
    res = _rand(B, sum(n), 64, seed=2).cuda().requires_grad_()
    back = A.bridge_merge(list(slabs), res)
    assert torch.equal(back, t + res)
    g = _rand(B, sum(n), 64, seed=3).cuda()
    back.backward(g)
    assert torch.equal(t.grad, g) and torch.equal(res.grad, g)


def test_scale_reduce_pack_forward_backward(cuda_lib):
    """Scale_reduce without its LayerNorm (MSTr.py:2225-2247) against the reference's op sequence (three strided convs on NCHW
    views, the `reshape(B, C, -1).permute(0, 2, 1)` re-reading, concatenation with the raw stage-4 tokens)."""
    from transception_b200 import autograd as A
    B, S, C = 2, 16, 64
    h = [S, S // 2, S // 4, S // 8]
    n = [h[0] ** 2, h[1] ** 2 * 2, h[2] ** 2 * 5, h[3] ** 2 * 8]
    x = _rand(B, sum(n), C, seed=1)
    ws = [0.05 * _rand(C * m, C * m, r, r, seed=10 + i) for i, (m, r) in enumerate(((1, 8), (2, 4), (5, 2)))]
    bs = [0.1 * _rand(C * m, seed=20 + i) for i, m in enumerate((1, 2, 5))]
    nred = (S // 8) ** 2 * 16
    dy = _rand(B, nred, C, seed=30)

    def ref(x, w0, b0, w1, b1, w2, b2):
        t0 = x[:, :n[0]].reshape(B, h[0], h[0], C).permute(0, 3, 1, 2)
        t1 = x[:, n[0]:n[0] + n[1]].reshape(B, h[1], h[1], C * 2).permute(0, 3, 1, 2)
        t2 = x[:, n[0] + n[1]:n[0] + n[1] + n[2]].reshape(B, h[2], h[2], C * 5).permute(0, 3, 1, 2)
        t3 = x[:, n[0] + n[1] + n[2]:]
        s0 = F.conv2d(t0, w0, b0, stride=8).reshape(B, C, -1).permute(0, 2, 1)
        s1 = F.conv2d(t1, w1, b1, stride=4).reshape(B, C, -1).permute(0, 2, 1)
        s2 = F.conv2d(t2, w2, b2, stride=2).reshape(B, C, -1).permute(0, 2, 1)
        return torch.cat([s0, s1, s2, t3], -2)
    r = [t.clone().requires_grad_() for t in (x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    yr = ref(*r)
    yr.backward(dy)
    g = [t.cuda().requires_grad_() for t in (x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    y = A.scale_reduce_pack(*g)
    _close(y, yr, 2e-3, "packed")                       # TF32 tensor-core GEMMs
    y.backward(dy.cuda())
    for name, a, b in zip(("dx", "dw0", "db0", "dw1", "db1", "dw2", "db2"), g, r):
        _close(a.grad, b.grad, 5e-3, name, floor=1e-4)
    first = [t.grad.clone() for t in g]
    for t in g:
        t.grad = None
    A.scale_reduce_pack(*g).backward(dy.cuda())
    assert all(torch.equal(a, t.grad) for a, t in zip(first, g))


@pytest.mark.parametrize("cin", [1, 3])
def test_patch_embed_conv_node(cuda_lib, cin):
    """OverlapPatchEmbeddings.proj (7x7 / 4, pad 3; MSTr.py:299-302); a 1-channel image stands for three equal planes (:2828-2829)."""
    from transception_b200 import autograd as A
    B, H, W = 2, 40, 36
    x = _rand(B, cin, H, W, seed=1)
    w, b = 0.1 * _rand(64, 3, 7, 7, seed=2), 0.1 * _rand(64, seed=3)
    xr = x.repeat(1, 3, 1, 1) if cin == 1 else x
    wr, br = w.clone().requires_grad_(), b.clone().requires_grad_()
    yr = F.conv2d(xr, wr, br, stride=4, padding=3)
    dy = _rand(*yr.shape, seed=4)
    yr.backward(dy)
    wg, bg = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = A.patch_embed_conv(x.cuda(), wg, bg)
    Ho, Wo = yr.shape[2], yr.shape[3]
    _close(y.view(B, Ho, Wo, 64).permute(0, 3, 1, 2), yr, 1e-5, "conv")
    y.backward(dy.permute(0, 2, 3, 1).reshape(B, Ho * Wo, 64).cuda())
    _close(wg.grad, wr.grad, 3e-3, "dw")                # TF32 weight-gradient GEMM
    _close(bg.grad, br.grad, 3e-3, "db")


@pytest.mark.parametrize("C", [64, 128, 256, 320, 512, 96])
def test_layernorm_dual(cuda_lib, C):
    from transception_b200 import ops
    x = (_rand(77, C, seed=1, scale=2.0) + 0.3).cuda()
    w, b = (1 + 0.2 * _rand(C, seed=2)).cuda(), (0.1 * _rand(C, seed=3)).cuda()
    y, y16 = ops.layernorm_dual(x, w, b, 1e-5)
    assert torch.equal(y, ops.layernorm(x, w, b, 1e-5))
    if C in ops.LN_DUAL_WIDTHS:
        assert y16 is not None and y16.dtype == torch.float16 and torch.equal(y16, y.half())
    else:
        assert y16 is None


def test_fan_out_sums_the_alias_gradients(cuda_lib):
    """autograd.FanOutFn: n aliases of one parameter; the gradient that reaches the parameter is the sum of the aliases' gradients
    in index order, from one kernel — equal to what AccumulateGrad's n - 1 additions give."""
    from transception_b200 import autograd as A
    w = _rand(48, 1, 7, 7, seed=1).cuda().requires_grad_()
    gs = [_rand(48, 1, 7, 7, seed=10 + i).cuda() for i in range(8)]
    al = A.fan_out(w, 8)
    assert all(a.data_ptr() == w.data_ptr() for a in al)
    torch.autograd.backward([a for i, a in enumerate(al) if i != 5], [g for i, g in enumerate(gs) if i != 5])     # alias 5 unused
    want = gs[0].clone()
    for i in (1, 2, 3, 4, 6, 7):
        want += gs[i]
    assert torch.equal(w.grad, want)
