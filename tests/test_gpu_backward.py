"""GPU parity of the training row's backward entries: gradients from the library's kernels
(through the autograd nodes of transception_b200/autograd.py -> C ABI) against torch autograd over the CPU oracle
(the reference's gradients ARE ATen autograd over its forward ops) on the same seeded weights and inputs.

Tolerance (stated): forward activations are fp16 and the gradient GEMMs run TF32 (operands read in place, fp32 accumulation),
so per tensor relative L2 error <= 1e-2 and cosine >= 0.999 (SURVEY §8d asks cosine >= 0.99); plain LayerNorm backward is fp32
arithmetic: <= 1e-4 relative.  Every result must be bit-identical run to run (no atomics).
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import mstr_oracle as O

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _pfloor(name, floor):
    """Absolute floor for a parameter gradient: bias gradients are summed on the tensor core in TF32 (the gradient operand is
    truncated to a 10-bit mantissa), so a bias whose true gradient cancels to zero (keys under a softmax over tokens, a conv
    bias in front of BatchNorm) is left with ~1e-5..1e-4 of the incoming gradient norm instead of fp32 rounding noise."""
    return floor * (100.0 if name.endswith("bias") else 1.0)


def _check(got, want, rel, what, floor=0.0):
    """relative L2 <= rel (+ an absolute floor for gradients that are analytically zero, e.g. the bias of the keys under a
    softmax over tokens, where both sides are rounding noise) and cosine >= 0.999"""
    got = got.float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.isfinite(got).all(), what + ": non-finite gradient"
    den = want.norm().item()
    err = (got - want).norm().item()
    assert err <= rel * den + floor, "%s: relative L2 %.3e > %.1e (|want| %.3e)" % (what, err / max(den, 1e-30), rel, den)
    if den > 100 * floor and den > 0:
        cos = F.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        assert cos >= 0.999, "%s: cosine %.5f" % (what, cos)
    return err / max(den, 1e-12)


@pytest.mark.parametrize("M,C,eps", [(300, 64, 1e-5), (97, 320, 1e-6), (50, 2048, 1e-5), (1, 128, 1e-5), (4099, 256, 1e-5)])
def test_layernorm_bwd(cuda_lib, M, C, eps):
    from transception_b200 import autograd as A
    x = _rand(M, C, seed=1, scale=2.0) + 0.3
    w = 1 + 0.2 * _rand(C, seed=2)
    b = 0.1 * _rand(C, seed=3)
    dy = _rand(M, C, seed=4)
    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    F.layer_norm(xr, (C,), wr, br, eps).backward(dy)
    xg, wg, bg = (t.cuda().requires_grad_() for t in (x, w, b))
    A.layernorm(xg, wg, bg, eps).backward(dy.cuda())
    _check(xg.grad, xr.grad, 1e-4, "LN dx")
    _check(wg.grad, wr.grad, 1e-4, "LN dw")
    _check(bg.grad, br.grad, 1e-4, "LN db")
    first = [t.grad.clone() for t in (xg, wg, bg)]
    for t in (xg, wg, bg):
        t.grad = None
    A.layernorm(xg, wg, bg, eps).backward(dy.cuda())
    for a, t in zip(first, (xg, wg, bg)):
        assert torch.equal(a, t.grad), "LayerNorm backward is not bit-reproducible"


@pytest.mark.parametrize("T,NL,KL,batch", [(256, 128, 64, 0), (64, 64, 64, 0), (100, 64, 64, 0), (777, 256, 64, 0), (5000, 320, 1280, 0),
                                           (50176, 256, 64, 0), (50176, 64, 256, 0), (6272, 16, 64, 0), (12544, 512, 128, 0),
                                           (784, 64, 64, 16), (3136, 128, 128, 16), (196, 320, 320, 16), (49, 512, 512, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_wgrad_mn_major(cuda_lib, T, NL, KL, batch, dtype):
    """out = alpha * a^T b over the token axis with both operands read in place as MN-major tcgen05 tiles (wgrad_tc.cu): every
    split plan (direct, one cluster, clusters + HBM level), bias sums, head mask, transposed copy.  Exact products of fp16 / bf16
    inputs accumulate in fp32, so those match fp64 to fp32 rounding (5e-6); fp32 inputs go through TF32 (10-bit mantissa): 2e-3.
    Bit-reproducible."""
    from transception_b200 import ops
    g = torch.Generator().manual_seed(T + NL)
    a = (torch.randn(((batch,) if batch else ()) + (T, NL), generator=g) * 0.5).to(dtype)
    b = torch.randn(((batch,) if batch else ()) + (T, KL), generator=g).to(dtype)
    want = torch.matmul(a.double().transpose(-1, -2), b.double()) * 0.25
    ch = NL // 8 if batch else 0
    if ch:
        want = want * (torch.arange(NL)[:, None] // ch == torch.arange(KL)[None, :] // ch).double()
    out, outT, db = ops.wgrad_mn(a.cuda(), b.cuda(), alpha=0.25, need_db=not batch, need_T=bool(batch), mask_ch=ch)
    tol = 2e-3 if dtype == torch.float32 else 5e-6
    assert (out.double().cpu() - want).norm() <= tol * want.norm()
    if batch:
        assert torch.equal(outT, out.transpose(-1, -2))
    else:
        wdb = a.double().sum(0) * 0.25
        assert (db.double().cpu() - wdb).norm() <= tol * wdb.norm()
    out2, _, db2 = ops.wgrad_mn(a.cuda(), b.cuda(), alpha=0.25, need_db=not batch, need_T=bool(batch), mask_ch=ch)
    assert torch.equal(out, out2) and (batch or torch.equal(db, db2)), "weight-gradient kernel is not bit-reproducible"


@pytest.mark.parametrize("M,N,K", [(1000, 256, 64), (777, 64, 256), (50, 2048, 512), (6272, 128, 512), (33, 320, 1280)])
@pytest.mark.parametrize("x16", [False, True])
def test_linear_bwd(cuda_lib, M, N, K, x16):
    from transception_b200 import ops
    x = _rand(M, K, seed=1)
    w = _rand(N, K, seed=2, scale=K ** -0.5)
    dy = _rand(M, N, seed=3, scale=1e-4)          # the magnitude a per-pixel loss gradient has: below fp16's normal range
    if x16:
        x = x.half().float()
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    br = torch.zeros(N, requires_grad=True)
    F.linear(xr, wr, br).backward(dy)
    xg = x.cuda().half() if x16 else x.cuda()
    dx, dw, db = ops.linear_bwd(xg, w.cuda(), dy.cuda())
    _check(dx, xr.grad, 2e-3, "linear dx")
    _check(dw, wr.grad, 2e-3, "linear dw")
    # the bias gradient rides on the tensor core (one TF32 MMA per k-step against a tile of ones) instead of an fp32 column sum
    _check(db, br.grad, 2e-3, "linear db")
    dx2, dw2, db2 = ops.linear_bwd(xg, w.cuda(), dy.cuda())
    assert torch.equal(dx, dx2) and torch.equal(dw, dw2) and torch.equal(db, db2)
    only = ops.linear_bwd(xg, w.cuda(), dy.cuda(), need_dx=False, need_db=False)
    assert only[0] is None and only[2] is None and torch.equal(only[1], dw)


def _mix_module(C, seed):
    from networks.MSTr import MixFFN_skip
    torch.manual_seed(seed)
    m = MixFFN_skip(C, 4 * C)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        m.norm1.weight.copy_(1 + 0.2 * torch.randn(4 * C, generator=g))
        m.norm1.bias.copy_(0.1 * torch.randn(4 * C, generator=g))
        for lin in (m.fc1, m.fc2):
            lin.bias.copy_(0.05 * torch.randn(lin.bias.shape, generator=g))
        m.dwconv.dwconv.bias.copy_(0.05 * torch.randn(4 * C, generator=g))
    return m


@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 64), (1, 28, 28, 128), (2, 7, 7, 320), (2, 7, 7, 512), (1, 56, 56, 64), (3, 5, 9, 64)])
def test_mixffn_skip_backward(cuda_lib, B, H, W, C):
    """MixFFN_skip (MSTr.py:58-61) forward + backward through the drop-in module in autograd mode."""
    m = _mix_module(C, seed=C + H)
    sd = {k: v.clone().requires_grad_() for k, v in m.state_dict().items()}
    x = _rand(B, H * W, C, seed=7)
    dy = _rand(B, H * W, C, seed=8, scale=1e-3)
    xr = x.clone().requires_grad_()
    sdp = {"m." + k: v for k, v in sd.items()}
    want = O.mixffn_skip(sdp, "m", xr, H, W)
    want.backward(dy)
    mg = m.cuda().train()
    xg = x.cuda().requires_grad_()
    got = mg(xg, H, W)
    assert (got.float().cpu() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    got.backward(dy.cuda())
    _check(xg.grad, xr.grad, 1e-2, "mixffn dx")
    live = ["fc1.weight", "fc1.bias", "dwconv.dwconv.weight", "dwconv.dwconv.bias", "norm1.weight", "norm1.bias",
            "fc2.weight", "fc2.bias"]
    params = dict(mg.named_parameters())
    for k in live:
        _check(params[k].grad, sdp["m." + k].grad, 1e-2, "mixffn d " + k)
    # the dead parameters (norm2 / norm3, MSTr.py:56-57) stay without gradient, as in the reference
    for k in ("norm2.weight", "norm2.bias", "norm3.weight", "norm3.bias"):
        assert params[k].grad is None and sdp["m." + k].grad is None
    first = {k: params[k].grad.clone() for k in live}
    dx1 = xg.grad.clone()
    for p in params.values():
        p.grad = None
    xg.grad = None
    mg(xg, H, W).backward(dy.cuda())
    assert torch.equal(dx1, xg.grad)
    for k in live:
        assert torch.equal(first[k], params[k].grad), k + ": not bit-reproducible"


def test_mixffn_skip_backward_tracks_weight_updates(cuda_lib):
    """An optimizer step changes the weights in place: the next forward/backward must see the new values."""
    m = _mix_module(64, seed=3).cuda().train()
    x = _rand(2, 49, 64, seed=1).cuda()
    dy = _rand(2, 49, 64, seed=2).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    m(x, 7, 7).backward(dy)
    opt.step()
    opt.zero_grad(set_to_none=True)
    y1 = m(x, 7, 7)
    y1.backward(dy)
    sd = {"m." + k: v.detach().cpu().clone().requires_grad_() for k, v in m.state_dict().items()}
    want = O.mixffn_skip(sd, "m", x.cpu(), 7, 7)
    want.backward(dy.cpu())
    assert (y1.detach().cpu() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    _check(m.fc2.weight.grad, sd["m.fc2.weight"].grad, 1e-2, "fc2.weight after step")
    _check(m.fc1.weight.grad, sd["m.fc1.weight"].grad, 1e-2, "fc1.weight after step")


def _randomise(net, seed=5):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight.copy_(1 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and m.bias is not None:
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    return net


def _grad_parity(module, prefix, oracle_fn, x, dy, rel=1e-2, dead=()):
    """module on CUDA in autograd mode vs torch autograd over the CPU oracle; returns the parameter names checked."""
    sd = {prefix + "." + k: v.clone().requires_grad_() for k, v in module.state_dict().items()}
    # shared modules (one ConvPosEnc / ConvRelPosEnc per MHCAEncoder) appear under several state_dict keys: the oracle reads
    # each alias as its own leaf, so the reference gradient of the shared parameter is the sum over its aliases
    alias = {}
    for k, v in module.state_dict(keep_vars=True).items():
        alias.setdefault(id(v), []).append(k)
    alias = {ks[0]: ks for ks in alias.values()}
    xr = x.clone().requires_grad_()
    want = oracle_fn(sd, xr)
    want.backward(dy)
    mg = module.cuda().train()
    xg = x.cuda().requires_grad_()
    got = mg_call(mg, xg)
    assert (got.float().cpu() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    got.backward(dy.cuda())
    floor = 1e-6 * dy.norm().item()
    _check(xg.grad, xr.grad, rel, "dx")
    params = dict(mg.named_parameters())
    checked = []
    for k, p in params.items():
        refs = [sd[prefix + "." + a].grad for a in alias[k] if sd[prefix + "." + a].grad is not None]
        ref = sum(refs) if refs else None
        if ref is None:
            assert p.grad is None, k + ": gradient where the reference has none"
            continue
        assert p.grad is not None, k + ": no gradient"
        _check(p.grad, ref, rel, "d " + k, _pfloor(k, floor))
        checked.append(k)
    first = {k: params[k].grad.clone() for k in checked}
    dx1 = xg.grad.clone()
    for p in params.values():
        p.grad = None
    xg.grad = None
    mg_call(mg, xg).backward(dy.cuda())
    assert torch.equal(dx1, xg.grad), "dx not bit-reproducible"
    for k in checked:
        assert torch.equal(first[k], params[k].grad), k + ": not bit-reproducible"
    return checked


def mg_call(mod, x):
    return mod(x, *mod._call_hw) if getattr(mod, "_call_hw", None) else mod(x)


@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 64), (1, 28, 28, 128), (2, 7, 7, 320), (1, 56, 56, 64), (2, 6, 11, 64)])
def test_efficient_attention_backward(cuda_lib, B, H, W, C):
    """EfficientAttention (MSTr.py:106-143), NCHW in / out as the reference module."""
    from networks.MSTr import EfficientAttention
    torch.manual_seed(C + H)
    m = _randomise(EfficientAttention(C, C, C, 1))
    x = _rand(B, C, H, W, seed=3)
    dy = _rand(B, C, H, W, seed=4, scale=1e-3)
    checked = _grad_parity(m, "a", lambda sd, xr: O.efficient_attention(sd, "a", xr), x, dy)
    assert len(checked) == 8


@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 64), (1, 28, 28, 128), (2, 7, 7, 320), (1, 56, 56, 64)])
def test_efficient_block_backward(cuda_lib, B, H, W, C):
    """EfficientTransformerBlock (MSTr.py:164-173): LN -> attention -> LN -> Mix-FFN with both residuals."""
    from networks.MSTr import EfficientTransformerBlock
    torch.manual_seed(C + W)
    m = _randomise(EfficientTransformerBlock(C, C, C, head_count=1, token_mlp="mix_skip"))
    m._call_hw = (H, W)
    x = _rand(B, H * W, C, seed=5)
    dy = _rand(B, H * W, C, seed=6, scale=1e-3)
    checked = _grad_parity(m, "b", lambda sd, xr: O.efficient_block(sd, "b", xr, H, W), x, dy)
    assert len(checked) == 4 + 8 + 8          # two LayerNorms, attention, the live Mix-FFN parameters


@pytest.mark.parametrize("scale,H,W,C", [(2, 7, 7, 512), (2, 14, 14, 320), (4, 14, 14, 64)])
def test_patch_expand_backward(cuda_lib, scale, H, W, C):
    """PatchExpand / FinalPatchExpand_X4 (MSTr.py:184-201, :212-227)."""
    from networks.MSTr import FinalPatchExpand_X4, PatchExpand
    torch.manual_seed(C)
    m = _randomise(PatchExpand((H, W), C, 2) if scale == 2 else FinalPatchExpand_X4((H, W), C, 4))
    x = _rand(2, H * W, C, seed=5)
    cout = C // 2 if scale == 2 else C
    dy = _rand(2, H * W * scale * scale, cout, seed=6, scale=1e-3)
    checked = _grad_parity(m, "u", lambda sd, xr: O.patch_expand(sd, "u", xr, H, W, scale), x, dy)
    assert len(checked) == 3


@pytest.mark.parametrize("is_last", [False, True])
def test_decoder_layer_backward(cuda_lib, is_last):
    """MyDecoderLayer (MSTr.py:273-290) with a skip map: concat Linear, two efficient blocks, expand (+ class head)."""
    from networks.MSTr import MyDecoderLayer
    torch.manual_seed(11)
    h = w = 14
    io = [32, 64, 64, 64] if is_last else [144, 128, 128, 128]
    m = _randomise(MyDecoderLayer((h, w), io, 1, "mix_skip", n_class=9, is_last=is_last))
    c1 = io[1]
    c2 = io[0] * (4 if is_last else 2) - c1
    x1 = _rand(2, h * w, c1, seed=1)
    x2 = _rand(2, h, w, c2, seed=2)
    sd = {"d." + k: v.clone().requires_grad_() for k, v in m.state_dict().items()}
    x1r, x2r = x1.clone().requires_grad_(), x2.clone().requires_grad_()
    want = O.decoder_layer(sd, "d", x1r, x2r, is_last=is_last)
    dy = _rand(*want.shape, seed=3, scale=1e-3)
    want.backward(dy)
    mg = m.cuda().train()
    x1g, x2g = x1.cuda().requires_grad_(), x2.cuda().requires_grad_()
    got = mg(x1g, x2g)
    assert got.shape == want.shape
    assert (got.float().cpu() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    got.backward(dy.cuda())
    floor = 1e-6 * dy.norm().item()
    _check(x1g.grad, x1r.grad, 1e-2, "dx1")
    _check(x2g.grad, x2r.grad, 1e-2, "dx2")
    n = 0
    for k, p in mg.named_parameters():
        ref = sd["d." + k].grad
        if ref is None:
            assert p.grad is None, k
            continue
        _check(p.grad, ref, 1e-2, "d " + k, _pfloor(k, floor))
        n += 1
    assert n == 2 + 2 * 20 + 3 + (2 if is_last else 0)


@pytest.mark.parametrize("B,H,W,C,add", [(2, 14, 14, 64, True), (1, 7, 9, 128, False), (2, 7, 7, 320, True)])
def test_dwconv_tokens_backward(cuda_lib, B, H, W, C, add):
    """ConvPosEnc (MSTr.py:744-752, with the skip) and DWConv (:26-31)."""
    from networks.MSTr import ConvPosEnc, DWConv
    torch.manual_seed(C)
    m = _randomise(ConvPosEnc(C) if add else DWConv(C))
    m._call_hw = ((H, W),) if add else (H, W)
    x = _rand(B, H * W, C, seed=2)
    dy = _rand(B, H * W, C, seed=3, scale=1e-3)

    def oracle(sd, xr):
        if add:
            return O.conv_pos_enc(sd, "c", xr, H, W)
        return O.map_to_tokens(O.conv(sd, "c.dwconv", O.tokens_to_map(xr, H, W), 1, 1, C))
    assert len(_grad_parity(m, "c", oracle, x, dy, rel=1e-4)) == 2


def _mhca_encoder(C, layers):
    from networks.MSTr import MHCAEncoder
    torch.manual_seed(C)
    return _randomise(MHCAEncoder(C, num_layers=layers, num_heads=8, mlp_ratio=4, drop_path_list=[0.0] * layers))


@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 64), (1, 14, 14, 128), (2, 7, 7, 320), (1, 28, 28, 64), (2, 5, 8, 64)])
def test_factor_att_backward(cuda_lib, B, H, W, C):
    """FactorAtt_ConvRelPosEnc (MSTr.py:852-886) with the 3/5/7 conv relative position encoding (:801-823)."""
    enc = _mhca_encoder(C, 1)
    m = enc.MHCA_layers[0].factoratt_crpe
    m._call_hw = ((H, W),)
    x = _rand(B, H * W, C, seed=2)
    dy = _rand(B, H * W, C, seed=3, scale=1e-3)
    checked = _grad_parity(m, "f", lambda sd, xr: O.factor_att(sd, "f", xr, H, W), x, dy)
    assert len(checked) == 2 + 6 + 2


@pytest.mark.parametrize("B,H,W,C,L", [(2, 14, 14, 64, 2), (1, 14, 14, 128, 1), (2, 7, 7, 320, 2)])
def test_mhca_encoder_backward(cuda_lib, B, H, W, C, L):
    """MHCAEncoder (MSTr.py:981-993): L MHCABlocks sharing one ConvPosEnc / ConvRelPosEnc, NCHW output."""
    m = _mhca_encoder(C, L)
    m._call_hw = ((H, W),)
    x = _rand(B, H * W, C, seed=2)
    dy = _rand(B, C, H, W, seed=3, scale=1e-3)
    _grad_parity(m, "e", lambda sd, xr: O.mhca_encoder(sd, "e", xr, H, W, L), x, dy)


@pytest.mark.parametrize("B,Nq,Nk", [(2, 300, 64), (1, 6076, 784), (2, 1000, 200)])
def test_attention_core_backward(cuda_lib, B, Nq, Nk):
    """softmax(q k^T / 8) v of M_EfficientSelfAtten (MSTr.py:2281-2285): flash forward, GEMM-recompute backward."""
    from transception_b200 import autograd as A
    q = _rand(B, Nq, 64, seed=1)
    kv = _rand(B, Nk, 128, seed=2)
    dy = _rand(B, Nq, 64, seed=3, scale=1e-3)
    qr, kvr = q.clone().requires_grad_(), kv.clone().requires_grad_()
    want = ((qr @ kvr[..., :64].transpose(1, 2)) * 0.125).softmax(-1) @ kvr[..., 64:]
    want.backward(dy)
    qg, kvg = q.cuda().requires_grad_(), kv.cuda().requires_grad_()
    got = A.attn_core(qg, kvg, 0.125)
    assert (got.cpu() - want.detach()).abs().max().item() <= 2e-2
    got.backward(dy.cuda())
    _check(qg.grad, qr.grad, 1e-2, "attn dq")
    _check(kvg.grad, kvr.grad, 1e-2, "attn dkv")
    a, b = qg.grad.clone(), kvg.grad.clone()
    qg.grad = kvg.grad = None
    A.attn_core(qg, kvg, 0.125).backward(dy.cuda())
    assert torch.equal(a, qg.grad) and torch.equal(b, kvg.grad)


def _bridge_layer(ch_att, seed=21):
    from networks.MSTr import BridgLayer_4
    torch.manual_seed(seed)
    return _randomise(BridgLayer_4(64, 1, [1, 2, 4, 8], ch_att))


@pytest.mark.parametrize("ch_att", [True, False])
@pytest.mark.parametrize("S", [56, 32])
def test_bridge_attention_backward(cuda_lib, ch_att, S):
    """M_EfficientChannelAtten (MSTr.py:2309-2353) / M_EfficientSelfAtten + Scale_reduce (:2267-2292, :2225-2249)."""
    m = _bridge_layer(ch_att).attn
    ntok = S * S * 31 // 16
    x = _rand(1, ntok, 64, seed=4)
    dy = _rand(1, ntok, 64, seed=5, scale=1e-3)
    fn = O.bridge_channel_atten if ch_att else O.bridge_self_atten
    checked = _grad_parity(m, "a", lambda sd, xr: fn(sd, "a", xr), x, dy)
    assert len(checked) == (8 if ch_att else 6 + 8)      # the channel attention's scale_reduce is dead (MSTr.py:2306-2307)


@pytest.mark.parametrize("ch_att", [True, False])
def test_bridge_layer_backward(cuda_lib, ch_att):
    """BridgLayer_4 (MSTr.py:2373-2409) on the 224x224 token buffer [B, 6076, 64]."""
    m = _bridge_layer(ch_att)
    x = _rand(2, 6076, 64, seed=6)
    dy = _rand(2, 6076, 64, seed=7, scale=1e-3)
    _grad_parity(m, "l", lambda sd, xr: O.bridge_layer(sd, "l", xr, ch_att), x, dy)


def test_bridge_block_backward(cuda_lib):
    """BridgeBlock_4 (MSTr.py:2422-2442): four maps in, four maps out, gradients w.r.t. every map and parameter."""
    from networks.MSTr import BridgeBlock_4
    torch.manual_seed(3)
    m = _randomise(BridgeBlock_4(64, 1, [1, 2, 4, 8], [True, False, False, False]))
    shapes = [(64, 56), (128, 28), (320, 14), (512, 7)]
    maps = [_rand(1, c, h, h, seed=10 + i) for i, (c, h) in enumerate(shapes)]
    dys = [_rand(1, c, h, h, seed=20 + i, scale=1e-3) for i, (c, h) in enumerate(shapes)]
    sd = {"b." + k: v.clone().requires_grad_() for k, v in m.state_dict().items()}
    mr = [t.clone().requires_grad_() for t in maps]
    want = O.bridge_block(sd, "b", mr)
    torch.autograd.backward(want, dys)
    mg = m.cuda().train()
    xg = [t.cuda().requires_grad_() for t in maps]
    got = mg(xg)
    for g_, w_ in zip(got, want):
        assert (g_.float().cpu() - w_.detach()).abs().max().item() <= 2e-2 * max(1.0, w_.abs().max().item())
    torch.autograd.backward(got, [d.cuda() for d in dys])
    floor = 1e-6 * sum(d.norm().item() for d in dys)
    for i in range(4):
        _check(xg[i].grad, mr[i].grad, 1e-2, "d map %d" % i)
    n = 0
    for k, p in mg.named_parameters():
        ref = sd["b." + k].grad
        if ref is None:
            assert p.grad is None, k
            continue
        _check(p.grad, ref, 1e-2, "d " + k, _pfloor(k, floor))
        n += 1
    assert n > 150


@pytest.fixture
def bn_train():
    """oracle BatchNorm in train mode (batch statistics) for the duration of a test"""
    O.BN_TRAIN = True
    yield
    O.BN_TRAIN = False


@pytest.mark.parametrize("M,C,act", [(1568, 64, 2), (392, 128, 0), (98, 320, 2), (112, 16, 4), (50, 80, 4)])
def test_bn_act_train(cuda_lib, M, C, act):
    """BatchNorm2d (batch statistics, running update) + Hardswish / silu_swish: forward, backward, running statistics."""
    from transception_b200 import autograd as A
    bn = torch.nn.BatchNorm2d(C)
    g = torch.Generator().manual_seed(C)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g))
        bn.bias.copy_(0.1 * torch.randn(C, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(C, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(C, generator=g))
    ref = torch.nn.BatchNorm2d(C)
    ref.load_state_dict(bn.state_dict())
    x = _rand(M, C, seed=1, scale=1.5) + 0.2
    dy = _rand(M, C, seed=2, scale=1e-3)
    fact = {0: lambda t: t, 2: F.hardswish, 4: O.silu_swish}[act]
    xr = x.clone().requires_grad_()
    want = fact(ref.train()(xr.t().reshape(1, C, M, 1))).reshape(C, M).t()
    want.backward(dy)
    bn = bn.cuda().train()
    xg = x.cuda().requires_grad_()
    got = A.bn_act(xg, bn, act)
    assert (got.cpu() - want.detach()).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())
    got.backward(dy.cuda())
    _check(xg.grad, xr.grad, 1e-3, "bn dx", floor=1e-9)
    _check(bn.weight.grad, ref.weight.grad, 1e-3, "bn dw", floor=1e-9)
    _check(bn.bias.grad, ref.bias.grad, 1e-3, "bn db", floor=1e-9)
    assert (bn.running_mean.cpu() - ref.running_mean).abs().max().item() <= 1e-5
    assert (bn.running_var.cpu() - ref.running_var).abs().max().item() <= 1e-5
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 1


@pytest.mark.parametrize("B,H,W,C,stride", [(2, 14, 14, 64, 2), (2, 14, 14, 64, 1), (1, 7, 9, 128, 2), (2, 7, 7, 320, 1)])
def test_dwconv3x3_nhwc_backward(cuda_lib, B, H, W, C, stride):
    from transception_b200 import autograd as A
    x = _rand(B, H, W, C, seed=1)
    w = _rand(C, 1, 3, 3, seed=2, scale=0.3)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    want = F.conv2d(xr.permute(0, 3, 1, 2), wr, None, stride, 1, 1, C).permute(0, 2, 3, 1)
    dy = _rand(*want.shape, seed=3, scale=1e-3)
    want.backward(dy)
    xg, wg = x.cuda().requires_grad_(), w.cuda().requires_grad_()
    got = A.dwconv3x3_nhwc(xg, wg, stride)
    assert got.shape == want.shape and (got.cpu() - want.detach()).abs().max().item() <= 1e-5
    got.backward(dy.cuda())
    _check(xg.grad, xr.grad, 1e-4, "dw3 dx")
    _check(wg.grad, wr.grad, 1e-4, "dw3 dw")


def test_ripm_stage_backward(cuda_lib, bn_train):
    """Patch_Embed_stage (MSTr.py:725-732): three chained dw3x3 -> 1x1 -> BatchNorm(train) -> Hardswish, first one stride 2."""
    from networks.MSTr import Patch_Embed_stage
    torch.manual_seed(2)
    m = _randomise(Patch_Embed_stage(64, num_path=3, isPool=True))
    x = _rand(2, 64, 28, 28, seed=1)
    sd = {"r." + k: v.clone().requires_grad_() if v.is_floating_point() else v.clone() for k, v in m.state_dict().items()}
    xr = x.clone().requires_grad_()
    want = O.patch_embed_stage(sd, "r", xr)
    dys = [_rand(*t.shape, seed=5 + i, scale=1e-3) for i, t in enumerate(want)]
    torch.autograd.backward(want, dys)
    mg = m.cuda().train()
    xg = x.cuda().requires_grad_()
    got = mg(xg)
    for a, b in zip(got, want):
        assert (a.float().cpu() - b.detach()).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item())
    torch.autograd.backward(got, [d.cuda() for d in dys])
    # Hardswish has a discontinuous derivative at z = -3 and z = 3 (jumps of 0.5): after three chained TF32 1x1 convs the
    # pre-activations differ from the fp32 oracle by ~1e-3, so a handful of the 392 pixels of a channel sit on the other side
    # of a kink and change that channel's gradient by O(dy); the single-block cases (fp32-exact forward) hold 1e-3 / 1e-2.
    _check(xg.grad, xr.grad, 5e-2, "ripm dx")
    for k, p in mg.named_parameters():
        _check(p.grad, sd["r." + k].grad, 5e-2, "ripm d " + k, floor=1e-8)


def _module_parity_train(m, prefix, oracle_fn, inputs, rel=1e-2):
    """generic: module on CUDA in train mode vs oracle autograd; inputs = list of tensors passed positionally"""
    sd = {prefix + "." + k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
    alias = {}
    for k, v in m.state_dict(keep_vars=True).items():
        alias.setdefault(id(v), []).append(k)
    alias = {ks[0]: ks for ks in alias.values()}
    ir = [t.clone().requires_grad_() for t in inputs]
    want = oracle_fn(sd, ir)
    dy = _rand(*want.shape, seed=77, scale=1e-3)
    want.backward(dy)
    mg = m.cuda().train()
    ig = [t.cuda().requires_grad_() for t in inputs]
    got = mg(ig) if getattr(m, "_list_input", False) else mg(*ig)
    assert (got.float().cpu() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    got.backward(dy.cuda())
    floor = 1e-6 * dy.norm().item()
    for i, (a, b) in enumerate(zip(ig, ir)):
        _check(a.grad, b.grad, rel, "d input %d" % i, floor)
    n = 0
    for k, p in mg.named_parameters():
        refs = [sd[prefix + "." + a].grad for a in alias[k] if sd[prefix + "." + a].grad is not None]
        if not refs:
            assert p.grad is None, k
            continue
        _check(p.grad, sum(refs), rel, "d " + k, _pfloor(k, floor))
        n += 1
    return n


def test_resblock_backward(cuda_lib, bn_train):
    from networks.MSTr import ResBlock
    torch.manual_seed(4)
    m = _randomise(ResBlock(128, 128))
    n = _module_parity_train(m, "r", lambda sd, i: O.resblock(sd, "r", i[0]), [_rand(2, 128, 14, 14, seed=1)])
    assert n == 3 + 6


def test_coordatt_backward(cuda_lib, bn_train):
    """IFF: CoordAtt (MSTr.py:1322-1348) over 4C concatenated channels."""
    from networks.MSTr import CoordAtt
    torch.manual_seed(5)
    m = _randomise(CoordAtt(256, 128, reduction=16))
    n = _module_parity_train(m, "c", lambda sd, i: O.coord_att(sd, "c", i[0]), [_rand(2, 256, 14, 12, seed=1)])
    assert n == 2 + 2 + 2 + 2 + 2


def test_mhca_stage_backward(cuda_lib, bn_train):
    """MHCA_stage (MSTr.py:1412-1441): ResBlock + three MHCA encoders -> concat -> CoordAtt."""
    from networks.MSTr import MHCA_stage
    torch.manual_seed(6)
    m = _randomise(MHCA_stage(64, 128, num_layers=2, num_heads=8, mlp_ratio=4, num_path=3, drop_path_list=[0.0, 0.0], concat='coord'))
    m._list_input = True
    ins = [_rand(2, 64, 14, 14, seed=10 + i) for i in range(3)]
    _module_parity_train(m, "s", lambda sd, i: O.mhca_stage(sd, "s", i, 2), ins)


def test_stem_backward(cuda_lib):
    """OverlapPatchEmbeddings (MSTr.py:299-304): conv 7x7 / 4 + LayerNorm; gradients of the conv and LayerNorm parameters."""
    from networks.MSTr import OverlapPatchEmbeddings
    torch.manual_seed(7)
    m = _randomise(OverlapPatchEmbeddings(64, 7, 4, 3, 3, 64))
    x = _rand(2, 3, 64, 64, seed=1)
    sd = {"p." + k: v.clone().requires_grad_() for k, v in m.state_dict().items()}
    want = O.patch_embed(sd, "p", x)[0]
    dy = _rand(*want.shape, seed=2, scale=1e-3)
    want.backward(dy)
    mg = m.cuda().train()
    got = mg(x.cuda())[0]
    assert (got.float().cpu() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    got.backward(dy.cuda())
    for k, p in mg.named_parameters():
        _check(p.grad, sd["p." + k].grad, 1e-2, "stem d " + k)


def test_whole_model_train_step(cuda_lib, bn_train):
    """MSTransception.forward in train mode + 0.4 CE + 0.6 Dice (trainer.py:139-146) + backward: loss and every parameter
    gradient against torch autograd over the CPU oracle (bs2, same seeded weights); the set of parameters without gradient
    (dead parameters) must be identical."""
    from networks.MSTr import MSTransception
    from oracle import loss_oracle as LO
    from transception_b200.losses import CeDiceLoss
    torch.manual_seed(1234)
    net = _randomise(MSTransception(num_classes=9))
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in net.state_dict().items()}
    alias = {}
    for k, v in net.state_dict(keep_vars=True).items():
        alias.setdefault(id(v), []).append(k)
    alias = {ks[0]: ks for ks in alias.values()}
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 224, 224, generator=gen) * 2 - 1
    labels = torch.randint(0, 9, (2, 224, 224), generator=gen)
    logits_ref = O.forward(sd, x)
    loss_ref = LO.ce_dice(logits_ref, labels, 9)[0]
    loss_ref.backward()
    mg = net.cuda().train()
    logits = mg(x.cuda())
    err = (logits.float().cpu() - logits_ref.detach()).abs().max().item()
    assert err <= 5e-2, err
    loss = CeDiceLoss(9)(logits, labels.cuda())
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    loss.backward()
    worst, worst_rel, n_grad, n_none, e2, r2 = 1.0, 0.0, 0, 0, 0.0, 0.0
    for k, p in mg.named_parameters():
        refs = [sd[a].grad for a in alias[k] if sd[a].grad is not None]
        if not refs:
            assert p.grad is None, k + ": gradient where the reference has none"
            n_none += 1
            continue
        assert p.grad is not None, k + ": no gradient"
        ref = sum(refs)
        got = p.grad.float().cpu()
        assert torch.isfinite(got).all(), k
        den = ref.norm().item()
        err = (got - ref).norm().item()
        e2 += err * err
        r2 += den * den
        if den > 1e-7:
            cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
            worst = min(worst, cos)
            worst_rel = max(worst_rel, err / den)
            assert cos >= 0.99, "%s: cosine %.4f (|ref| %.3e)" % (k, cos, den)
            # per tensor relative L2: fp16 / TF32 tensor-core operands against fp32 autograd
            assert err <= 0.1 * den, "%s: relative L2 %.3e (|ref| %.3e)" % (k, err / den, den)
        n_grad += 1
    total_rel = (e2 / r2) ** 0.5
    print("train step: loss %.6f (oracle %.6f), %d parameters with gradient (worst cosine %.5f, worst relative L2 %.3e, "
          "relative L2 of the whole gradient %.3e), %d dead" % (loss.item(), loss_ref.item(), n_grad, worst, worst_rel, total_rel, n_none))
    assert total_rel <= 2e-2, "relative L2 of the whole gradient %.3e" % total_rel
    assert n_grad > 1000 and n_none > 100


def test_whole_model_train_step_other_geometry(cuda_lib, bn_train):
    """BASELINE config 5 geometry (256x256; the stock reference cannot run it, SURVEY §0 defect 4): the train step of
    MSTransception(image_size=256, num_classes=2) against autograd over the generalised oracle (bridge_geometry reduces to
    the reference constants at 224, where it is pinned)."""
    from networks.MSTr import MSTransception
    from oracle import loss_oracle as LO
    from transception_b200.losses import CeDiceLoss
    torch.manual_seed(7)
    net = _randomise(MSTransception(num_classes=2, image_size=256))
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in net.state_dict().items()}
    alias = {}
    for k, v in net.state_dict(keep_vars=True).items():
        alias.setdefault(id(v), []).append(k)
    alias = {ks[0]: ks for ks in alias.values()}
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 256, 256, generator=gen) * 2 - 1
    labels = torch.randint(0, 2, (1, 256, 256), generator=gen)
    loss_ref = LO.ce_dice(O.forward(sd, x), labels, 2)[0]
    loss_ref.backward()
    mg = net.cuda().train()
    loss = CeDiceLoss(2)(mg(x.cuda()), labels.cuda())
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    loss.backward()
    n = 0
    for k, p in mg.named_parameters():
        refs = [sd[a].grad for a in alias[k] if sd[a].grad is not None]
        if not refs:
            assert p.grad is None, k
            continue
        ref = sum(refs)
        if ref.norm().item() > 1e-7:
            cos = F.cosine_similarity(p.grad.float().cpu().flatten(), ref.flatten(), dim=0).item()
            assert cos >= 0.99, "%s: cosine %.4f" % (k, cos)
        n += 1
    assert n > 1000


def test_train_step_graph_matches_eager(cuda_lib):
    """runtime.TrainStepGraph: the captured step (forward + loss + backward + clip + SGD as one CUDA graph) updates the weights
    exactly like the same steps launched eagerly (kernels are deterministic, so the comparison is bit-exact)."""
    from networks.MSTr import MSTransception
    from transception_b200.losses import CeDiceLoss
    from transception_b200.runtime import TrainStepGraph
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand(2, 1, 224, 224, generator=gen) * 2 - 1).cuda()
    labels = torch.randint(0, 9, (2, 224, 224), generator=gen).cuda()

    def make():
        torch.manual_seed(1234)
        net = MSTransception(num_classes=9).cuda().train()
        return net, torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)

    net_a, opt_a = make()
    # 2 eager warm-up steps + the one real step PyTorch's capture recipe runs on a side stream before recording
    runner = TrainStepGraph(net_a, CeDiceLoss(9), opt_a, batch=2, warmup=2, max_norm=5.0, sample=(x, labels))
    assert runner.steps_done == 3
    losses = [float(runner.step(x, labels)) for _ in range(2)]                                      # + 2 replayed steps
    net_b, opt_b = make()
    crit = CeDiceLoss(9)
    eager = []
    for _ in range(5):
        opt_b.zero_grad(set_to_none=True)
        loss = crit(net_b(x), labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net_b.parameters(), max_norm=5.0, norm_type=2)
        opt_b.step()
        eager.append(float(loss.detach()))
    assert losses == eager[3:], (losses, eager)
    assert eager[-1] < eager[0]
    for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        assert torch.equal(a, b), k + ": graph-replayed training diverged from eager training"


def test_fused_sgd_matches_torch_sgd(cuda_lib):
    """optim.FusedSGD (multi-tensor kernels, device-side lr and clip coefficient) against clip_grad_norm_ + torch.optim.SGD
    (trainer.py:125,148) on tensors of ragged sizes, over several steps with a changing learning rate."""
    from transception_b200.optim import FusedSGD
    g = torch.Generator().manual_seed(3)
    shapes = [(7,), (64, 64), (4097,), (320, 1280), (3, 5, 7), (1,), (12289,)]
    pa = [torch.randn(s, generator=g).cuda().requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    unused = torch.randn(5).cuda().requires_grad_()          # never receives a gradient: must be left alone
    oa = FusedSGD(pa + [unused], lr=0.05, momentum=0.9, weight_decay=1e-4, max_norm=0.5)
    ob = torch.optim.SGD(pb, lr=0.05, momentum=0.9, weight_decay=1e-4)
    for it, lr in enumerate((0.05, 0.05, 0.01, 0.2)):
        grads = [torch.randn(s, generator=g).cuda() * (0.1 + it) for s in shapes]
        for p, q, gr in zip(pa, pb, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        for grp in oa.param_groups + ob.param_groups:
            grp["lr"] = lr
        oa.step()
        norm = torch.nn.utils.clip_grad_norm_(pb, max_norm=0.5, norm_type=2)
        ob.step()
        assert abs(float(oa.norm_coef[0]) - float(norm)) <= 1e-5 * float(norm)
        for p, q in zip(pa, pb):
            assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item()), (it, tuple(p.shape))
    assert unused.grad is None and "momentum_buffer" not in oa.state.get(unused, {})
    for p, q in zip(pa, pb):
        assert (oa.state[p]["momentum_buffer"] - ob.state[q]["momentum_buffer"]).abs().max().item() <= 1e-5


def test_train_step_graph_fused_sgd(cuda_lib):
    """TrainStepGraph + FusedSGD: graph-replayed training equals the same steps launched eagerly bit for bit, tracks
    torch.optim.SGD training closely, keeps the forward's fp16 weight copies current without conversion launches, and follows a
    per-iteration learning-rate schedule (trainer.py:151-153) without re-capture."""
    from networks.MSTr import MSTransception
    from transception_b200 import ops
    from transception_b200.losses import CeDiceLoss
    from transception_b200.optim import FusedSGD
    from transception_b200.runtime import TrainStepGraph
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand(2, 1, 224, 224, generator=gen) * 2 - 1).cuda()
    labels = torch.randint(0, 9, (2, 224, 224), generator=gen).cuda()
    lrs = [0.05, 0.05, 0.05, 0.03, 0.02, 0.0]

    def make(fused):
        torch.manual_seed(1234)
        net = MSTransception(num_classes=9).cuda().train()
        cls = FusedSGD if fused else torch.optim.SGD
        return net, cls(net.parameters(), lr=lrs[0], momentum=0.9, weight_decay=1e-4)

    net_a, opt_a = make(True)
    with torch.no_grad():
        net_a.eval()(x)              # the inference path prepares fp16 copies the training path never touches: they must not go stale
    runner = TrainStepGraph(net_a, CeDiceLoss(9), opt_a, batch=2, warmup=2, sample=(x, labels))
    assert runner.steps_done == 3
    losses = []
    for lr in lrs[3:]:
        for grp in opt_a.param_groups:
            grp["lr"] = lr                                   # what trainer.py:151-153 does every iteration
        if lr == 0.0:
            before = {k: v.clone() for k, v in net_a.named_parameters()}
        losses.append(float(runner.step(x, labels)))
    # lr = 0 in the last step: momentum still moves, the weights must not
    for k, v in net_a.named_parameters():
        assert torch.equal(v, before[k]), k + ": the update graph did not pick up the new learning rate"
    crit = CeDiceLoss(9)

    def eager(net, opt):
        out = []
        for lr in lrs:
            for grp in opt.param_groups:
                grp["lr"] = lr
            opt.zero_grad(set_to_none=True)
            loss = crit(net(x), labels)
            loss.backward()
            opt.step()
            out.append(float(loss.detach()))
        return out
    net_b, opt_b = make(True)
    eb = eager(net_b, opt_b)
    assert losses == eb[3:], (losses, eb)
    for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        assert torch.equal(a, b), k + ": graph-replayed FusedSGD training diverged from eager FusedSGD training"
    net_c, opt_c = make(False)
    ec = eager(net_c, opt_c)
    assert all(abs(a - b) <= 2e-3 * abs(b) for a, b in zip(eb, ec)), (eb, ec)
    assert ec[-2] < ec[0]
    # the prepared fp16 copies follow the raw-pointer updates: eval logits of the graph-trained net == a fresh eval of its weights
    net_a.eval()
    with torch.no_grad():
        y1 = net_a(x)
        ops.invalidate_prepared(net_a)
        y2 = net_a(x)
    assert torch.equal(y1, y2), "stale fp16 weight copies after FusedSGD updates"


@pytest.mark.parametrize("B,Nq,Nk,big", [(2, 300, 64, False), (1, 128, 128, False), (3, 777, 200, False), (2, 6076, 784, False),
                                         (1, 1000, 784, True)])
def test_flash_attention_backward(cuda_lib, B, Nq, Nk, big):
    """The tcgen05 flash backward of softmax(q k^T / 8) v (MSTr.py:2281-2285) against torch autograd in fp64: ragged query / kv
    tile edges, the production shape, large scores (|s| up to ~20), gradients of the magnitude a per-pixel loss produces.
    fp16 operands with fp32 accumulation and a per-call dO scale: relative L2 <= 5e-3.  Bit-reproducible."""
    from transception_b200 import ops
    g = torch.Generator().manual_seed(Nq + Nk)
    q = torch.randn(B, Nq, 64, generator=g) * (2.5 if big else 1.0)
    kv = torch.randn(B, Nk, 128, generator=g) * (2.5 if big else 1.0)
    dout = torch.randn(B, Nq, 64, generator=g) * 1e-6
    qr, kvr = q.double().requires_grad_(), kv.double().requires_grad_()
    k, v = kvr[..., :64], kvr[..., 64:]
    want = torch.softmax(qr @ k.transpose(1, 2) * 0.125, dim=-1) @ v
    want.backward(dout.double())
    out, lse = ops.flash_attn_train(q.cuda(), kv.cuda(), 0.125)
    assert (out.cpu().double() - want.detach()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    ref_lse = torch.logsumexp(qr.detach() @ k.detach().transpose(1, 2) * 0.125, dim=-1) * 1.4426950408889634
    assert (lse.cpu().double() - ref_lse).abs().max().item() <= 2e-2
    dq, dkv = ops.flash_attn_bwd(q.cuda(), kv.cuda(), out, lse, dout.cuda(), 0.125)
    for got, ref, name in ((dq, qr.grad, "dq"), (dkv[..., :64], kvr.grad[..., :64], "dk"), (dkv[..., 64:], kvr.grad[..., 64:], "dv")):
        err = (got.cpu().double() - ref).norm().item() / ref.norm().item()
        assert torch.isfinite(got).all() and err <= 5e-3, "%s: relative L2 %.3e" % (name, err)
    dq2, dkv2 = ops.flash_attn_bwd(q.cuda(), kv.cuda(), out, lse, dout.cuda(), 0.125)
    assert torch.equal(dq, dq2) and torch.equal(dkv, dkv2), "flash backward is not bit-reproducible"


def test_split_backward_and_bucket_parts(cuda_lib):
    """The pieces of the overlapped data-parallel step (runtime.TrainStepGraph, world > 1) on one GPU: the backward cut between
    encoder stages 2 and 3 (MSViT._grad_cut) followed by the second piece gives bit-identical gradients to one backward pass; after
    the first piece exactly the parameters in front of the cut are still without gradient; FusedSGD.set_bucket_tail orders the flat
    bucket as [behind the cut | in front of it] and gather_grads(0) / gather_grads(1) fill exactly those two parts."""
    from networks.MSTr import MSTransception
    from transception_b200.losses import CeDiceLoss
    from transception_b200.optim import FusedSGD
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand(2, 1, 224, 224, generator=gen) * 2 - 1).cuda()
    labels = torch.randint(0, 9, (2, 224, 224), generator=gen).cuda()
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).cuda().train()
    crit = CeDiceLoss(9)
    crit(net(x), labels).backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    bb = net.backbone
    early = {id(p) for name in bb.EARLY_MODULES for p in getattr(bb, name).parameters()}
    bb._grad_cut = []
    try:
        crit(net(x), labels).backward()
        cut = list(bb._grad_cut)
    finally:
        bb._grad_cut = None
    assert len(cut) == 2
    for k, p in net.named_parameters():
        if k in ref:
            assert (p.grad is None) == (id(p) in early), k
    torch.autograd.backward([o for o, _ in cut], [leaf.grad for _, leaf in cut])
    for k, p in net.named_parameters():
        assert (p.grad is not None) == (k in ref), k
        if k in ref:
            assert torch.equal(p.grad, ref[k]), k + ": split backward differs from the single pass"
    # bucket order and the two gathers
    opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    opt.set_bucket_tail([p for p in net.parameters() if id(p) in early])
    head = opt.gather_grads(0)
    tab = opt._tables[0]
    used = tab["used"]
    n_head = sum(1 for p in used if id(p) not in early)
    assert all(id(p) not in early for p in used[:n_head]) and all(id(p) in early for p in used[n_head:])
    assert 0 < n_head < len(used) and head.numel() == tab["head_elems"] and head.numel() > 0.9 * tab["total"]
    offs = tab["offs"].tolist()
    flat = tab["flat"]
    assert float(flat[tab["head_elems"]:].abs().sum()) == 0.0                       # the tail has not been gathered yet
    tail = opt.gather_grads(1)
    assert head.numel() + tail.numel() == tab["total"]
    for i in (0, n_head - 1, n_head, len(used) - 1):
        p = used[i]
        assert torch.equal(flat[offs[i]:offs[i] + p.numel()], p.grad.flatten()), i
    whole = flat.clone()
    flat.zero_()
    opt.gather_grads()
    assert torch.equal(flat, whole)
