"""GPU parity of the networks/Transception.py variant: drop-in modules -> C ABI -> sm_100a kernels
against the CPU oracle (oracle/transception_oracle.py) and against the golden fixtures recorded from the real reference.

The variant has one back end (fp16 tensor-core operands and fp16 intermediates, fp32 accumulation and fp32 residual
stream), so the tensor-core bar of SURVEY §8d applies: per-op max-abs <= 2e-2 * max(1, absmax); logits max-abs <= 5e-2,
mean-abs <= 5e-3, argmax Dice >= 0.99."""
import os

import pytest
import torch

from oracle import fixtures as FX
from oracle import transception_oracle as TO
from oracle.cases import flatten_out
from oracle.cases_transception import BY_NAME, CASE_NAMES, seeded_model

pytestmark = pytest.mark.gpu
TC_TOL = 2e-2
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transception_golden.pt")


@pytest.fixture(scope="module")
def model(cuda_lib):
    net = seeded_model(perturb=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    return net.cuda(), sd


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _close(got, want, tol, what):
    got = got.float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.isfinite(got).all(), what + ": non-finite output"
    err = (got - want).abs().max().item()
    ref = max(1.0, want.abs().max().item())
    assert err <= tol * ref, "%s: max-abs err %.3e > %.1e * %.3g" % (what, err, tol, ref)
    return err


def _dice(a, b, ncls):
    d = []
    for c in range(ncls):
        x, y = a == c, b == c
        den = x.sum().item() + y.sum().item()
        if den:
            d.append(2.0 * (x & y).sum().item() / den)
    return min(d)


def _cuda_args(args):
    return [a.cuda() if isinstance(a, torch.Tensor) else a for a in args]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_module_matches_oracle_and_golden(model, golden, name):
    net, sd = model
    _, path, mk, fn = BY_NAME[name]
    args = mk()
    with torch.no_grad():
        want = flatten_out(fn(sd, path, *args))
        got = flatten_out(net.get_submodule(path)(*_cuda_args(args)))
    g = golden[name]
    for i, (a, b) in enumerate(zip(got, want)):
        _close(a, b, TC_TOL, "%s[%d]" % (name, i))
        sub = FX.subsample(a.float().cpu())
        ref = max(1.0, float(g["stats"][i][2]))
        assert (sub - g["sub"][i]).abs().max().item() <= TC_TOL * ref, name + " vs golden"


@pytest.mark.parametrize("stage,C,hw,bs", [(2, 64, 56, 2), (3, 128, 28, 3), (4, 320, 14, 2)])
def test_inception_stage(model, stage, C, hw, bs):
    """dual patch merging -> 2 fuse blocks -> stage norm -> nearest upsample + concat + 1x1 conv"""
    net, sd = model
    x = FX.rand(bs, C, hw, hw, seed=60 + stage)
    with torch.no_grad():
        want = TO.fuse_stage(sd, 'backbone', x, stage)
        got = net.backbone.stage(x.cuda().permute(0, 2, 3, 1).contiguous(), stage).permute(0, 3, 1, 2)
    _close(got, want, TC_TOL * 2, "inception stage %d" % stage)


@pytest.mark.parametrize("cin,bs", [(1, 2), (3, 1)])
def test_whole_model(model, golden, cin, bs):
    net, sd = model
    x = FX.image(bs, cin, seed=0)
    with torch.no_grad():
        want = TO.forward(sd, x, return_all=True)
        maps = net.backbone(x.cuda())
        for i in range(4):
            _close(maps[i], want['enc'][i], TC_TOL * 3, "encoder map %d" % i)
        got = net(x.cuda()).float().cpu()
    w = want['logits']
    assert got.shape == w.shape == (bs, 9, 224, 224)
    err, mean = (got - w).abs().max().item(), (got - w).abs().mean().item()
    dice = _dice(got.argmax(1), w.argmax(1), 9)
    print("Transception logits max-abs %.3e mean-abs %.3e min-class Dice %.5f" % (err, mean, dice))
    assert err <= 5e-2 and mean <= 5e-3 and dice >= 0.99
    g = golden["model_c%d" % cin]
    assert (got[:, :, ::4, ::4] - g["logits_sub"]).abs().max().item() <= 5e-2


@pytest.mark.parametrize("stage,C,hw,bs", [(2, 64, 56, 2), (4, 320, 14, 3)])
def test_inception_stage_sk_fusion(model, stage, C, hw, bs):
    """same stage with the selective-kernel fusion tail (concat='sk': SK_Block, Transception.py:328-358)"""
    net, sd = model
    x = FX.rand(bs, C, hw, hw, seed=70 + stage)
    net.backbone.concat = 'sk'
    try:
        with torch.no_grad():
            want = TO.fuse_stage(sd, 'backbone', x, stage, concat='sk')
            got = net.backbone.stage(x.cuda().permute(0, 2, 3, 1).contiguous(), stage).permute(0, 3, 1, 2)
    finally:
        net.backbone.concat = 'original'
    _close(got, want, TC_TOL * 2, "inception stage %d (sk)" % stage)


def test_whole_model_sk_fusion(cuda_lib, golden):
    net = seeded_model(perturb=True, concat='sk')
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    x = FX.image(1, 1, seed=0)
    with torch.no_grad():
        want = TO.forward(sd, x, concat='sk')
        got = net(x.cuda()).float().cpu()
    err, mean = (got - want).abs().max().item(), (got - want).abs().mean().item()
    print("Transception(concat='sk') logits max-abs %.3e mean-abs %.3e" % (err, mean))
    assert err <= 5e-2 and mean <= 5e-3
    assert (got[:, :, ::4, ::4] - golden["model_sk"]["logits_sub"]).abs().max().item() <= 5e-2


def test_forward_is_bit_reproducible_and_graph_capturable(model):
    from transception_b200.runtime import GraphRunner
    net, _ = model
    x = FX.image(2, 1, seed=3).cuda()
    with torch.no_grad():
        a = net(x).clone()
        b = net(x).clone()
    assert torch.equal(a, b)
    runner = GraphRunner(net, 2, 1, 224, device="cuda")
    runner.x.copy_(x)
    runner.graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(runner.y, a)


def test_unprepared_path_fails_loudly(cuda_lib):
    """The variant has no TF32 / FFMA back end: with the fp16 pipeline switched off the call raises, it does not fall back."""
    from transception_b200 import ops
    net = seeded_model(perturb=False).cuda()
    ops.set_flag("f16_pipeline", 0)
    try:
        with pytest.raises(RuntimeError):
            with torch.no_grad():
                net.backbone.block2[0].attn(torch.zeros(1, 1460, 128, device="cuda"))
    finally:
        ops.set_flag("f16_pipeline", 1)
