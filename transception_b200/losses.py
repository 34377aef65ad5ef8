"""Training loss of the reference (``trainer.py:122-123, 141-143``; ``utils.py:11-47``) on the sm_100a library.

``DiceLoss`` mirrors ``utils.DiceLoss`` (same constructor and ``forward(inputs, target, weight=None, softmax=False)``);
``CeDiceLoss`` is the fused form of the three lines ``loss_ce = ce_loss(outputs, label.long()); loss_dice =
dice_loss(outputs, label, softmax=True); loss = 0.4 * loss_ce + 0.6 * loss_dice``.  Both return a 0-dim device tensor with an
autograd node whose backward is the library's gradient kernel, and never synchronise with the host (the reference reads one
``.item()`` per class per step).  ``class_wise_dice`` of the last call is kept as a device tensor.
"""
import torch
import torch.nn as nn

from . import ops


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, w_ce, w_dice, class_w, softmax):
        out, ws = ops.seg_loss_fwd(logits, labels, w_ce, w_dice, class_w, softmax)
        ctx.save_for_backward(logits, labels, ws)
        ctx.cfg = (w_ce, w_dice, class_w, softmax)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, grad_loss, _grad_out):
        logits, labels, ws = ctx.saved_tensors
        w_ce, w_dice, class_w, softmax = ctx.cfg
        d = ops.seg_loss_bwd(logits, labels, ws, grad_loss, w_ce, w_dice, class_w, softmax)
        return d, None, None, None, None, None


class DiceLoss(nn.Module):
    """reference utils.py:11-47"""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes
        self.class_wise_dice = None

    def forward(self, inputs, target, weight=None, softmax=False):
        if inputs.shape[1] != self.n_classes:
            raise AssertionError('predict {} & target {} shape do not match'.format(inputs.size(), target.size()))
        loss, out = _SegLossFn.apply(inputs, target, 0.0, 1.0, None if weight is None else tuple(weight), bool(softmax))
        self.class_wise_dice = out[4:]
        return loss


class CeDiceLoss(nn.Module):
    """0.4 * CrossEntropyLoss + 0.6 * DiceLoss(softmax=True) in one pass (reference trainer.py:141-143)."""

    def __init__(self, n_classes, w_ce=0.4, w_dice=0.6):
        super().__init__()
        self.n_classes, self.w_ce, self.w_dice = n_classes, float(w_ce), float(w_dice)
        self.loss_ce = self.loss_dice = self.class_wise_dice = None

    def forward(self, outputs, label):
        if outputs.shape[1] != self.n_classes:
            raise AssertionError('predict {} & target {} shape do not match'.format(outputs.size(), label.size()))
        loss, out = _SegLossFn.apply(outputs, label, self.w_ce, self.w_dice, None, True)
        self.loss_ce, self.loss_dice, self.class_wise_dice = out[1], out[2], out[4:]
        return loss
