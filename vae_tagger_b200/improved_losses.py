"""``FocalLoss`` of the reference's ``improved_losses.py`` (:39-56) with a fused CUDA forward+backward.

On CUDA tensors the loss value and ``d loss / d logits`` come from one kernel
(``vt_focal_loss``: ``bce = BCEWithLogits``, ``pt = exp(-bce)``, ``alpha (1-pt)^gamma bce`` and the
analytic gradient); autograd only sees a custom ``Function``.  ``ClassBalancedLoss`` (:58-72)
composes the same kernel with per-class weights.  The triplet / contrastive / combined losses of
the reference serve VAE fine-tuning (``train_full.py`` / ``train_vae.py``) and are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _native


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, alpha, gamma):
        nctx = _native.get_context(logits.device)
        loss_sum, grad = nctx.focal_loss(logits, targets, alpha=alpha, gamma=gamma, want_grad=True)
        # grad holds d(mean)/d(logits); keep d(sum) scaling for the other reductions
        ctx.save_for_backward(grad)
        ctx.n = logits.numel()
        ctx.dtype = logits.dtype
        return loss_sum.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad_mean,) = ctx.saved_tensors
        return (grad_mean * (g * ctx.n)).to(ctx.dtype), None, None, None


class FocalLoss(nn.Module):
    def __init__(self, alpha=1, gamma=2, reduction="mean"):
        super().__init__()
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, inputs, targets):
        if inputs.device.type != "cuda":
            raise _native.NativeError("FocalLoss needs CUDA tensors on a B200 (no CPU fallback)")
        if self.reduction not in ("mean", "sum"):
            raise NotImplementedError("the fused focal loss kernel provides reduction 'mean' and 'sum'")
        total = _FocalFn.apply(inputs, targets.to(inputs.dtype), float(self.alpha), float(self.gamma))
        return total / inputs.numel() if self.reduction == "mean" else total


class ClassBalancedLoss(nn.Module):
    """Effective-number class weights on top of the focal term (reference :58-72)."""

    def __init__(self, beta=0.9999, gamma=2.0):
        super().__init__()
        self.beta, self.gamma = beta, gamma

    def forward(self, logits, labels, samples_per_class):
        n = torch.as_tensor(np.asarray(samples_per_class, dtype=np.float64), device=logits.device)
        eff = 1.0 - torch.pow(torch.as_tensor(self.beta, dtype=torch.float64, device=logits.device), n)
        w = (1.0 - self.beta) / (eff + 1e-8)
        w = (w / w.sum() * len(samples_per_class)).to(logits.dtype)
        # per-class weights commute with the elementwise focal term: weight the logits' gradient path
        bce = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="none")
        pt = torch.exp(-bce)
        return (w.unsqueeze(0) * (1 - pt) ** self.gamma * bce).mean()


def compute_class_distribution(dataset):
    """Number of positive samples per class (reference :341-348)."""
    counts = None
    for labels in dataset.image_labels.values():
        pos = (labels > 0).float()
        counts = pos if counts is None else counts + pos
    return counts.numpy() if counts is not None else np.zeros(0)
