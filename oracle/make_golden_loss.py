#!/usr/bin/env python
"""Generate tests/golden/loss_golden.pt from the REAL reference loss (authoring container only) — TEST INFRASTRUCTURE.

    python oracle/make_golden_loss.py

Imports /root/reference/utils.py (with stub modules for its unused, uninstalled imports medpy / SimpleITK) and evaluates
``DiceLoss`` + ``torch.nn.CrossEntropyLoss`` exactly as /root/reference/trainer.py:122-123, 141-143 does, on seeded
logits / labels, storing the loss terms and (subsampled) autograd gradients."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import fixtures as FX  # noqa: E402

REF = os.environ.get("TRANSCEPTION_REF", "/root/reference")
CASES = [("bs2_9c_224", 2, 9, 224, 3.0, 101), ("bs3_4c_64", 3, 4, 64, 1.0, 102), ("bs1_1c_32", 1, 1, 32, 1.0, 103),
         ("bs2_16c_48", 2, 16, 48, 8.0, 104)]


def case_inputs(B, K, S, scale, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, K, S, S, generator=g) * scale
    labels = torch.randint(0, K, (B, S, S), generator=g).float()       # the dataset hands float labels to trainer.py:137
    return logits, labels


def load_reference_utils():
    for name in ("medpy", "SimpleITK"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.metric = types.ModuleType(name + ".metric")
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("_tcx_ref_utils", os.path.join(REF, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    U = load_reference_utils()
    golden = {"meta": {"torch": torch.__version__, "reference_commit": "0c7ee13"}}
    for name, B, K, S, scale, seed in CASES:
        logits, labels = case_inputs(B, K, S, scale, seed)
        x = logits.clone().requires_grad_(True)
        ce_loss, dice_loss = torch.nn.CrossEntropyLoss(), U.DiceLoss(K)
        loss_ce = ce_loss(x, labels[:].long())
        loss_dice = dice_loss(x, labels, softmax=True)
        loss = 0.4 * loss_ce + 0.6 * loss_dice
        loss.backward()
        # DiceLoss alone on probabilities with class weights (utils.py:38-39)
        w = [0.5 + 0.1 * i for i in range(K)]
        pr = torch.softmax(logits, 1).clone().requires_grad_(True)
        ld = dice_loss(pr, labels, weight=w, softmax=False)
        ld.backward()
        golden[name] = {"loss": loss.detach(), "ce": loss_ce.detach(), "dice": loss_dice.detach(),
                        "grad_sub": FX.subsample(x.grad), "grad_stats": FX.stats(x.grad),
                        "dice_w": ld.detach(), "dice_w_grad_sub": FX.subsample(pr.grad), "weights": w}
        print(name, loss.item(), loss_ce.item(), loss_dice.item(), ld.item())
    path = os.path.join(ROOT, "tests", "golden", "loss_golden.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
