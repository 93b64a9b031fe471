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


def _need_cuda(t, what):
    if t.device.type != "cuda":
        raise _native.NativeError(f"{what} needs CUDA tensors on a B200 (no CPU fallback)")


class _EmbedLossFn(torch.autograd.Function):
    """Triplet (kind 0) / contrastive (kind 1) loss over flattened latents: value and all input gradients from one
    ``vt_embed_loss`` call."""

    @staticmethod
    def forward(ctx, kind, margin, similarity, a, p, n, labels_a, labels_p):
        nctx = _native.get_context(a.device)
        loss, grads = nctx.embed_loss(kind, a, p, n, labels_a, labels_p, margin=margin, similarity=similarity)
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.has_n = grads[2] is not None
        ctx.dtypes = (a.dtype, p.dtype, None if n is None else n.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        ga, gp = saved[0] * g, saved[1] * g
        gn = saved[2] * g if ctx.has_n else None
        return (None, None, None, ga.to(ctx.dtypes[0]), gp.to(ctx.dtypes[1]),
                None if gn is None else gn.to(ctx.dtypes[2]), None, None)


class ContrastiveLoss(nn.Module):
    """Label-similarity contrastive loss (reference :6-37): Jaccard similarity of the two label sets > 0.3 pulls the
    embeddings together (``d^2``), otherwise pushes them ``margin`` apart; weighted by the similarity."""

    def __init__(self, margin=1.0, similarity_type="cosine"):
        super().__init__()
        self.margin, self.similarity_type = margin, similarity_type

    def forward(self, embedding1, embedding2, labels1, labels2):
        _need_cuda(embedding1, "ContrastiveLoss")
        return _EmbedLossFn.apply(1, float(self.margin), self.similarity_type, embedding1, embedding2, None,
                                  labels1, labels2)


class ImprovedTripletLoss(nn.Module):
    """Triplet loss with optional label-overlap weights (reference :74-109)."""

    def __init__(self, margin=1.0, similarity_type="cosine"):
        super().__init__()
        self.margin, self.similarity_type = margin, similarity_type

    def forward(self, anchor, positive, negative, anchor_labels=None, positive_labels=None):
        _need_cuda(anchor, "ImprovedTripletLoss")
        if anchor_labels is None or positive_labels is None:
            anchor_labels = positive_labels = None
        return _EmbedLossFn.apply(0, float(self.margin), self.similarity_type, anchor, positive, negative,
                                  anchor_labels, positive_labels)


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        loss, gx = _native.get_context(x.device).mse_loss(x, y)
        ctx.save_for_backward(gx)
        ctx.dtypes = (x.dtype, y.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (gx,) = ctx.saved_tensors
        gx = gx * g
        return gx.to(ctx.dtypes[0]), (-gx).to(ctx.dtypes[1]) if ctx.needs_input_grad[1] else None


def mse_loss(x, y):
    """``F.mse_loss(x, y)`` (mean) with the value and the gradient from one kernel pass."""
    _need_cuda(x, "mse_loss")
    return _MseFn.apply(x, y.expand_as(x))


class _AdaptiveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_w, losses, temperature):
        total, weights, glw = _native.get_context(log_w.device).adaptive_loss_weights(log_w, losses, temperature)
        ctx.save_for_backward(weights, glw)
        ctx.mark_non_differentiable(weights)
        return total.reshape(()), weights

    @staticmethod
    def backward(ctx, g, _gw):
        weights, glw = ctx.saved_tensors
        return glw * g, weights * g, None


class AdaptiveLossWeights(nn.Module):
    """Learnable softmax weighting of ``num_losses`` scalar losses (reference :111-125).  Returns
    ``(weighted_loss, weights)``; like the reference, the weights are ``softmax(log_weights / temperature)``."""

    def __init__(self, num_losses=4, temperature=1.0):
        super().__init__()
        self.num_losses, self.temperature = num_losses, temperature
        self.log_weights = nn.Parameter(torch.zeros(num_losses))

    def forward(self, losses):
        stacked = torch.stack([l.reshape(()) for l in losses]).to(self.log_weights.device)
        _need_cuda(stacked, "AdaptiveLossWeights")
        return _AdaptiveFn.apply(self.log_weights, stacked, float(self.temperature))


class SimplifiedCombinedLoss(nn.Module):
    """Semantic loss (triplet or contrastive) + classification loss (reference :127-232; same arguments, same
    result dictionary)."""

    def __init__(self, classification_weight=1.0, triplet_weight=0.5, contrastive_weight=0.0, use_focal_loss=True,
                 use_class_balanced=False, use_contrastive=False, focal_alpha=1.0, focal_gamma=2.0, triplet_margin=1.0,
                 contrastive_margin=1.0, similarity_type="cosine"):
        super().__init__()
        self.classification_weight, self.triplet_weight = classification_weight, triplet_weight
        self.contrastive_weight, self.use_contrastive = contrastive_weight, use_contrastive
        self.classification_loss_fn = FocalLoss(alpha=focal_alpha, gamma=focal_gamma) if use_focal_loss \
            else nn.BCEWithLogitsLoss()
        self.use_class_balanced = use_class_balanced
        if use_class_balanced:
            self.class_balanced_loss_fn = ClassBalancedLoss()
        if use_contrastive:
            self.contrastive_loss_fn = ContrastiveLoss(margin=contrastive_margin, similarity_type=similarity_type)
        else:
            self.triplet_loss_fn = ImprovedTripletLoss(margin=triplet_margin, similarity_type=similarity_type)

    def forward(self, z_a, z_p, z_n=None, classification_logits=None, classification_targets=None, anchor_labels=None,
                positive_labels=None, negative_labels=None, samples_per_class=None):
        loss_dict, total = {}, 0
        if self.use_contrastive and self.contrastive_weight > 0:
            c = self.contrastive_loss_fn(z_a.reshape(z_a.size(0), -1), z_p.reshape(z_p.size(0), -1), anchor_labels,
                                         positive_labels)
            total = total + self.contrastive_weight * c
            loss_dict["contrastive_loss"] = c
        elif self.triplet_weight > 0:
            t = self.triplet_loss_fn(z_a.reshape(z_a.size(0), -1), z_p.reshape(z_p.size(0), -1),
                                     z_n.reshape(z_n.size(0), -1), anchor_labels, positive_labels)
            total = total + self.triplet_weight * t
            loss_dict["triplet_loss"] = t
        if classification_logits is not None and classification_targets is not None:
            if self.use_class_balanced and samples_per_class is not None:
                cl = self.class_balanced_loss_fn(classification_logits, classification_targets, samples_per_class)
            else:
                cl = self.classification_loss_fn(classification_logits, classification_targets)
            total = total + self.classification_weight * cl
            loss_dict["classification_loss"] = cl
        loss_dict["total_loss"] = total
        loss_dict["weights"] = torch.tensor([self.contrastive_weight if self.use_contrastive else self.triplet_weight,
                                             self.classification_weight])
        return loss_dict


class CombinedLoss(nn.Module):
    """Reconstruction + log-stabilised KL + triplet + classification (reference :234-339), fixed or adaptive
    weights."""

    def __init__(self, reconstruction_weight=0.01, kl_weight=1e-2, triplet_weight=1.0, classification_weight=1.0,
                 use_focal_loss=True, use_class_balanced=False, use_adaptive_weights=False, focal_alpha=1.0,
                 focal_gamma=2.0, triplet_margin=1.0, similarity_type="cosine"):
        super().__init__()
        self.reconstruction_weight, self.kl_weight = reconstruction_weight, kl_weight
        self.triplet_weight, self.classification_weight = triplet_weight, classification_weight
        self.classification_loss_fn = FocalLoss(alpha=focal_alpha, gamma=focal_gamma) if use_focal_loss \
            else nn.BCEWithLogitsLoss()
        self.use_class_balanced = use_class_balanced
        if use_class_balanced:
            self.class_balanced_loss_fn = ClassBalancedLoss()
        self.triplet_loss_fn = ImprovedTripletLoss(margin=triplet_margin, similarity_type=similarity_type)
        self.use_adaptive_weights = use_adaptive_weights
        if use_adaptive_weights:
            self.adaptive_weights = AdaptiveLossWeights(num_losses=4)

    def forward(self, reconstruction, target_images, posterior_a, posterior_p, posterior_n, z_a, z_p, z_n,
                classification_logits, classification_targets, anchor_labels=None, positive_labels=None,
                samples_per_class=None):
        reconstruction_loss = mse_loss(reconstruction, target_images)
        kl_mean = ((posterior_a.kl() + posterior_p.kl() + posterior_n.kl()) / 3).mean()
        kl_loss = torch.log(1 + kl_mean / 10000)
        triplet_loss = self.triplet_loss_fn(z_a.reshape(z_a.size(0), -1), z_p.reshape(z_p.size(0), -1),
                                            z_n.reshape(z_n.size(0), -1), anchor_labels, positive_labels)
        if self.use_class_balanced and samples_per_class is not None:
            classification_loss = self.class_balanced_loss_fn(classification_logits, classification_targets,
                                                              samples_per_class)
        else:
            classification_loss = self.classification_loss_fn(classification_logits, classification_targets)
        loss_dict = {"reconstruction_loss": reconstruction_loss, "kl_loss": kl_loss, "triplet_loss": triplet_loss,
                     "classification_loss": classification_loss}
        if self.use_adaptive_weights:
            total, weights = self.adaptive_weights([reconstruction_loss, kl_loss, triplet_loss, classification_loss])
            loss_dict["adaptive_weights"] = weights
        else:
            total = (self.reconstruction_weight * reconstruction_loss + self.kl_weight * kl_loss +
                     self.triplet_weight * triplet_loss + self.classification_weight * classification_loss)
            loss_dict["weights"] = torch.tensor([self.reconstruction_weight, self.kl_weight, self.triplet_weight,
                                                 self.classification_weight])
        loss_dict["total_loss"] = total
        return loss_dict


def compute_class_distribution(dataset):
    """Number of positive samples per class (reference :341-348)."""
    counts = None
    for labels in dataset.image_labels.values():
        pos = (labels > 0).float()
        counts = pos if counts is None else counts + pos
    return counts.numpy() if counts is not None else np.zeros(0)
