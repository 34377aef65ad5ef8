"""CPU oracle of the reference training loss — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ``utils.DiceLoss`` (/root/reference/utils.py:11-47) and the loss lines of the training step
(/root/reference/trainer.py:122-123, 141-143) with plain PyTorch ops; gradients come from autograd.  Pinned against the real
``utils.DiceLoss`` / ``CrossEntropyLoss`` by ``oracle/make_golden_loss.py`` -> ``tests/golden/loss_golden.pt``."""
import torch
import torch.nn.functional as F


def one_hot(target, n_classes):                      # utils.py:16-22
    return torch.stack([(target == i) for i in range(n_classes)], dim=1).float()


def dice_loss(inputs, target, n_classes, weight=None, softmax=False):   # utils.py:24-47
    if softmax:
        inputs = torch.softmax(inputs, dim=1)
    t = one_hot(target, n_classes)
    weight = [1.0] * n_classes if weight is None else weight
    smooth = 1e-5
    loss, cls = 0.0, []
    for i in range(n_classes):
        s, y = inputs[:, i], t[:, i]
        d = 1 - (2 * torch.sum(s * y) + smooth) / (torch.sum(s * s) + torch.sum(y * y) + smooth)
        cls.append(1.0 - d.detach())
        loss = loss + d * weight[i]
    return loss / n_classes, torch.stack(cls)


def ce_dice(outputs, label, n_classes, w_ce=0.4, w_dice=0.6):       # trainer.py:141-143
    ce = F.cross_entropy(outputs, label.long())
    dice, cls = dice_loss(outputs, label, n_classes, softmax=True)
    return w_ce * ce + w_dice * dice, ce, dice, cls
