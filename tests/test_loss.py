"""Fused training loss.  CPU: the oracle (oracle/loss_oracle.py) against golden values recorded
from the REAL reference ``utils.DiceLoss`` + ``CrossEntropyLoss`` (oracle/make_golden_loss.py).  GPU: the library's fused
forward / backward kernels through the drop-in ``DiceLoss`` / ``CeDiceLoss`` modules against the oracle (fp32: relative
tolerance 2e-5 on the loss terms, 1e-4 of the gradient's absmax on gradients) and against the golden values."""
import os

import pytest
import torch

from oracle import fixtures as FX
from oracle import loss_oracle as LO
from oracle.make_golden_loss import CASES, case_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt")
BY = {c[0]: c for c in CASES}


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _oracle(name):
    _, B, K, S, scale, seed = BY[name]
    logits, labels = case_inputs(B, K, S, scale, seed)
    x = logits.clone().requires_grad_(True)
    loss, ce, dice, cls = LO.ce_dice(x, labels, K)
    loss.backward()
    return logits, labels, K, loss.detach(), ce.detach(), dice.detach(), cls, x.grad


@pytest.mark.parametrize("name", list(BY))
def test_oracle_matches_reference_loss(golden, name):
    g = golden[name]
    logits, labels, K, loss, ce, dice, cls, grad = _oracle(name)
    for got, key in ((loss, "loss"), (ce, "ce"), (dice, "dice")):
        assert abs(got.item() - g[key].item()) <= 1e-6 * max(1.0, abs(g[key].item())), key
    assert (FX.subsample(grad) - g["grad_sub"]).abs().max().item() <= 1e-7 + 1e-5 * float(g["grad_stats"][2])
    pr = torch.softmax(logits, 1).clone().requires_grad_(True)
    ld, _ = LO.dice_loss(pr, labels, K, weight=g["weights"], softmax=False)
    ld.backward()
    assert abs(ld.item() - g["dice_w"].item()) <= 1e-6
    assert (FX.subsample(pr.grad) - g["dice_w_grad_sub"]).abs().max().item() <= 1e-7


def test_mirror_surface():
    from transception_b200.losses import CeDiceLoss, DiceLoss
    d = DiceLoss(9)
    assert d.n_classes == 9
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.zeros(1, 9, 4, 4), torch.zeros(1, 4, 4), softmax=True)
    with pytest.raises(AssertionError, match="shape do not match"):
        CeDiceLoss(9)(torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(BY))
@pytest.mark.parametrize("label_dtype", [torch.float32, torch.int64, torch.uint8])
def test_fused_loss_matches_oracle(cuda_lib, golden, name, label_dtype):
    from transception_b200.losses import CeDiceLoss
    logits, labels, K, loss, ce, dice, cls, grad = _oracle(name)
    x = logits.cuda().requires_grad_(True)
    mod = CeDiceLoss(K)
    got = mod(x, labels.to(label_dtype).cuda())
    (got * 1.7).backward()                        # a non-unit upstream gradient exercises grad_out
    tol = lambda ref: 2e-5 * max(1.0, abs(ref))
    assert abs(got.item() - loss.item()) <= tol(loss.item())
    assert abs(mod.loss_ce.item() - ce.item()) <= tol(ce.item())
    assert abs(mod.loss_dice.item() - dice.item()) <= tol(dice.item())
    assert (mod.class_wise_dice.cpu() - cls).abs().max().item() <= 2e-5
    g = x.grad.cpu() / 1.7
    scale = max(grad.abs().max().item(), 1e-12)
    assert (g - grad).abs().max().item() <= 1e-4 * scale + 1e-9
    gg = golden[name]
    assert abs(got.item() - gg["loss"].item()) <= tol(gg["loss"].item())
    assert (FX.subsample(g) - gg["grad_sub"]).abs().max().item() <= 1e-4 * scale + 1e-9


@pytest.mark.gpu
def test_dice_loss_module_on_probabilities_with_weights(cuda_lib, golden):
    """DiceLoss(n)(probabilities, target, weight=w, softmax=False) — reference utils.py:33-47"""
    from transception_b200.losses import DiceLoss
    name = "bs3_4c_64"
    _, B, K, S, scale, seed = BY[name]
    logits, labels = case_inputs(B, K, S, scale, seed)
    w = golden[name]["weights"]
    pr = torch.softmax(logits, 1)
    p_ref = pr.clone().requires_grad_(True)
    want, _ = LO.dice_loss(p_ref, labels, K, weight=w, softmax=False)
    want.backward()
    p = pr.cuda().requires_grad_(True)
    got = DiceLoss(K)(p, labels.cuda(), weight=w, softmax=False)
    got.backward()
    assert abs(got.item() - want.item()) <= 2e-5 and abs(got.item() - golden[name]["dice_w"].item()) <= 2e-5
    assert (p.grad.cpu() - p_ref.grad).abs().max().item() <= 1e-4 * p_ref.grad.abs().max().item()


@pytest.mark.gpu
def test_fused_loss_is_deterministic_and_counts_bad_labels(cuda_lib):
    from transception_b200 import ops
    logits, labels = case_inputs(2, 9, 224, 3.0, 7)
    labels[0, 0, :5] = 12.0
    a, _ = ops.seg_loss_fwd(logits.cuda(), labels.cuda())
    b, _ = ops.seg_loss_fwd(logits.cuda(), labels.cuda())
    assert torch.equal(a, b) and a[3].item() == 5.0
