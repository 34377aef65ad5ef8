"""Golden vectors of the TRAINING row from the live reference: the shimmed
`networks.MSTr.MSTransception` in train mode (BatchNorm batch statistics) on a seeded bs2 input, the reference loss
0.4*CrossEntropy + 0.6*DiceLoss (trainer.py:141-143, utils.DiceLoss) and, per parameter, the gradient norm plus eight sampled
gradient values.  tests/test_oracle.py checks the oracle restatement (BN_TRAIN=True) + torch autograd against them.

Run in the authoring container (needs /root/reference):  python oracle/make_golden_train.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402


def main():
    ref = load_reference()
    for name in ("medpy", "medpy.metric", "SimpleITK"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["medpy"].metric = sys.modules["medpy.metric"]
    sys.path.insert(0, "/root/reference")
    from utils import DiceLoss  # the reference's own loss
    torch.manual_seed(1234)
    net = ref.MSTransception(num_classes=9).train()
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 224, 224, generator=gen) * 2 - 1
    labels = torch.randint(0, 9, (2, 224, 224), generator=gen)
    logits = net(x)
    loss_ce = torch.nn.CrossEntropyLoss()(logits, labels.long())
    loss_dice = DiceLoss(9)(logits, labels, softmax=True)
    loss = 0.4 * loss_ce + 0.6 * loss_dice
    loss.backward()
    grads = {}
    for k, p in net.named_parameters():
        if p.grad is None:
            grads[k] = None
            continue
        g = p.grad.flatten()
        idx = torch.linspace(0, g.numel() - 1, 8).long()
        grads[k] = (g.norm().item(), idx, g[idx].clone())
    out = {"loss": loss.item(), "ce": loss_ce.item(), "dice": loss_dice.item(),
           "logits_fingerprint": (logits.mean().item(), logits.std().item(), logits.abs().max().item()),
           "logits_sample": logits[:, :, ::37, ::41].detach().clone(), "grads": grads,
           "running_mean_after": {k: v.clone() for k, v in net.state_dict().items() if k.endswith("running_mean")}}
    path = os.path.join(ROOT, "tests", "golden", "train_golden.pt")
    torch.save(out, path)
    n = sum(v is not None for v in grads.values())
    print("wrote %s: loss %.6f, %d parameters with gradient, %d without" % (path, loss.item(), n, len(grads) - n))


if __name__ == "__main__":
    main()
