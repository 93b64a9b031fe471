"""``FocalLoss`` of the reference's ``improved_losses.py`` (:39-56) with a fused CUDA forward+backward.

On CUDA tensors the loss value and ``d loss / d logits`` come from one kernel
(``vt_focal_loss``: ``bce = BCEWithLogits``, ``pt = exp(-bce)``, ``alpha (1-pt)^gamma bce`` and the
analytic gradient); autograd only sees a custom ``Function``.  ``ClassBalancedLoss`` (:58-72) is the
reference's weighted BCE; the native training step applies its weights inside the same kernel.  The triplet / contrastive / combined losses of
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
            # any other reduction: the per-element focal tensor, as the reference returns it (:52-56)
            bce = torch.nn.functional.binary_cross_entropy_with_logits(inputs, targets.to(inputs.dtype),
                                                                       reduction="none")
            return self.alpha * (1 - torch.exp(-bce)) ** self.gamma * bce
        total = _FocalFn.apply(inputs, targets.to(inputs.dtype), float(self.alpha), float(self.gamma))
        return total / inputs.numel() if self.reduction == "mean" else total


def class_balanced_weights(samples_per_class, beta=0.9999):
    """Effective-number class weights exactly as the reference forms them (improved_losses.py:66-69): float64
    numpy, ``(1-beta)/(1-beta^n)`` normalised to sum to the number of classes."""
    spc = np.asarray(samples_per_class)
    effective_num = 1.0 - np.power(beta, spc)
    weights = (1.0 - beta) / effective_num
    return weights / weights.sum() * len(weights)


class ClassBalancedLoss(nn.Module):
    """Effective-number class weights on the binary cross entropy (reference :58-72; like the reference,
    ``gamma`` is accepted and unused).  On CUDA tensors the weighted loss and its gradient come from the fused
    kernel (``vt_focal_loss`` arithmetic with gamma = 0 and per-class weights)."""

    def __init__(self, beta=0.9999, gamma=2.0):
        super().__init__()
        self.beta, self.gamma = beta, gamma

    def forward(self, inputs, targets, samples_per_class):
        weights = torch.tensor(class_balanced_weights(samples_per_class, self.beta), dtype=torch.float32,
                               device=inputs.device)
        bce = torch.nn.functional.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
        return (bce * weights.unsqueeze(0)).mean()


class ClassBalancedCriterion(nn.Module):
    """``loss_fn(logits, labels)`` form of ``ClassBalancedLoss()(logits, labels, class_distribution)``
    (train_decoder.py:188-189) -- a module instead of a closure so that ``DecoderTrainer`` can hand the class
    weights to the native training step."""

    def __init__(self, class_distribution, beta=0.9999, gamma=2.0):
        super().__init__()
        self.loss = ClassBalancedLoss(beta, gamma)
        self.class_distribution = np.asarray(class_distribution)

    def weights(self):
        return class_balanced_weights(self.class_distribution, self.loss.beta)

    def forward(self, logits, labels):
        return self.loss(logits, labels, self.class_distribution)


def compute_class_distribution(dataset):
    """Number of positive samples per class (reference :341-348)."""
    counts = None
    for labels in dataset.image_labels.values():
        pos = (labels > 0).float()
        counts = pos if counts is None else counts + pos
    return counts.numpy() if counts is not None else np.zeros(0)
