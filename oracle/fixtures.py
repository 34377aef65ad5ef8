"""Shared seeded weights / inputs for the oracle, the golden generator and the tests — TEST INFRASTRUCTURE ONLY.

Everything here is deterministic in (seed, torch version): the weights are the drop-in mirror's same-seed
initialisation (identical to the reference's, proven by ``oracle/make_golden.py`` loading them strictly into
the reference model) with every affine / running-statistic tensor perturbed so that parity exercises it.
"""
import torch

MODEL_SEED = 1234
PERTURB_SEED = 5


def randomise(net, seed=PERTURB_SEED):
    """Make every LayerNorm/BatchNorm affine + running-stat tensor and every conv/linear bias non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.LayerNorm, torch.nn.BatchNorm2d)):
                m.weight.copy_(1 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and m.bias is not None:
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    return net


def seeded_model(num_classes=9, perturb=True):
    """The drop-in parameter mirror with the reference's same-seed init (+ optional perturbation), eval mode."""
    from transception_b200 import MSTransception
    torch.manual_seed(MODEL_SEED)
    net = MSTransception(num_classes=num_classes)
    if perturb:
        randomise(net)
    return net.eval()


def rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def image(bs, cin=1, seed=0, size=224):
    """Synthetic slice batch in [-1,1] (reference trainer.py:89-92 normalisation range)."""
    return torch.rand(bs, cin, size, size, generator=torch.Generator().manual_seed(seed)) * 2 - 1


def bridge_maps(seed, bs=2):
    return [rand(bs, c, s, s, seed=seed + i) for i, (c, s) in enumerate(((64, 56), (128, 28), (320, 14), (512, 7)))]


def subsample(t, limit=4096):
    """Deterministic strided subsample of a tensor to <= ~limit values (fixtures stay small)."""
    flat = t.detach().float().reshape(-1)
    step = max(1, flat.numel() // limit)
    return flat[::step].clone()


def stats(t):
    t = t.detach().double()
    return torch.tensor([t.mean().item(), t.std().item(), t.abs().max().item(), t.sum().item()], dtype=torch.float64)
