#!/usr/bin/env python
"""Fused CE+Dice loss forward+backward at the training-step shape (bs16, 9 classes, 224x224) on one GPU: CUDA-event time,
achieved bytes/s against the algorithmic traffic, and the same three reference lines (trainer.py:141-143, utils.py:11-47)
in eager PyTorch on the same GPU (with its per-class `.item()` syncs) for context.  Development aid, not the bench."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import loss_oracle as LO  # noqa: E402
from transception_b200.losses import CeDiceLoss  # noqa: E402


def main():
    B, K, S = 16, 9, 224
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn(B, K, S, S, generator=g) * 3).cuda()
    labels = torch.randint(0, K, (B, S, S), generator=g).float().cuda()
    mod = CeDiceLoss(K)

    def ours():
        x = logits.detach().requires_grad_(True)
        loss = mod(x, labels)
        loss.backward()
        return loss, x.grad

    def eager():
        x = logits.detach().requires_grad_(True)
        loss, ce, dice, cls = LO.ce_dice(x, labels, K)
        [c.item() for c in cls]                     # utils.py:45
        loss.backward()
        return loss, x.grad

    for name, fn in (("fused kernels", ours), ("eager PyTorch (reference lines)", eager)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50):
            fn()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / 50 * 1e3
        npix = B * S * S
        algo = npix * (K * 4 + 4) * 2 + npix * K * 4          # two reads of logits + labels, one gradient write
        print("%-34s %8.1f us / fwd+bwd  -> %.0f GB/s of algorithmic traffic (%.1f MB)" % (name, us, algo / us / 1e3, algo / 1e6))
    a, ga = ours()
    b, gb = eager()
    print("loss %.6f vs %.6f ; grad max-abs diff %.3e (absmax %.3e)" % (a.item(), b.item(), (ga - gb).abs().max().item(), gb.abs().max().item()))


if __name__ == "__main__":
    main()
