"""fp32 CPU restatement of the tag-decoder head, confidence sort, threshold and focal loss.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Pinned**: every function
here is checked against outputs of the reference's own classes (imported from
``/root/reference`` through a stub ``diffusers`` module by
``tests/golden/make_golden.py``) stored in ``tests/golden/head_golden.pt``.

Functional style: each function takes the reference ``state_dict`` (keys per
SURVEY.md Appendix B) so that one set of tensors drives oracle and CUDA path.

Reference lines followed:
  * SpatialAttention.forward ............... modules.py:36-47
  * feature_compress ....................... modules.py:377-382 (used :437)
  * MultiHeadSelfAttention.forward ......... modules.py:66-91
  * classifier ............................. modules.py:401-418 (used :462)
  * get_confidence ......................... modules.py:470-475
  * ClassificationDecoder (--no_attention) . modules.py:303-356
  * threshold loop ......................... infer_full.py:109-124
  * FocalLoss.forward ...................... improved_losses.py:47-56
Eval mode (dropout identity, BatchNorm running statistics) unless a function says otherwise; the
``*_train`` functions restate module.train() semantics (batch-statistics BatchNorm with running-buffer
update, Dropout with EXPLICIT keep-masks so that a checker can apply the masks of the path under test)
and are pinned by ``tests/golden/head_train_golden.pt`` (reference modules in train mode, gradients by
the reference's own autograd graph; ``tests/golden/make_train_golden.py``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def spatial_attention(sd: dict, x: torch.Tensor, prefix="spatial_attention.") -> torch.Tensor:
    w1 = sd[prefix + "channel_att.0.weight"]  # [C/8, C, 1, 1]
    w2 = sd[prefix + "channel_att.2.weight"]  # [C, C/8, 1, 1]
    w7 = sd[prefix + "spatial_att.0.weight"]  # [1, 2, 7, 7]

    def mlp(v):
        return F.conv2d(F.relu(F.conv2d(v, w1)), w2)

    avg = x.mean(dim=(2, 3), keepdim=True)
    mx = x.amax(dim=(2, 3), keepdim=True)
    x = x * torch.sigmoid(mlp(avg) + mlp(mx))
    m = torch.cat([x.mean(dim=1, keepdim=True), x.amax(dim=1, keepdim=True)], dim=1)
    return x * torch.sigmoid(F.conv2d(m, w7, padding=3))


def feature_compress(sd: dict, x: torch.Tensor, prefix="feature_compress.") -> torch.Tensor:
    y = F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], padding=1)
    y = F.batch_norm(
        y,
        sd[prefix + "1.running_mean"],
        sd[prefix + "1.running_var"],
        sd[prefix + "1.weight"],
        sd[prefix + "1.bias"],
        training=False,
        eps=1e-5,
    )
    return F.adaptive_avg_pool2d(F.relu(y), (8, 8))


def self_attention(sd: dict, x: torch.Tensor, heads: int = 8, prefix="self_attention_post.") -> torch.Tensor:
    b, c, h, w = x.shape
    n, hd = h * w, c // heads
    xf = x.reshape(b, c, n).transpose(1, 2)
    t = F.layer_norm(xf, (c,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps=1e-5)

    def proj(name):
        y = F.linear(t, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"])
        return y.reshape(b, n, heads, hd).transpose(1, 2)

    q, k, v = proj("q_proj"), proj("k_proj"), proj("v_proj")
    p = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(b, n, c)
    o = F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"]) + xf
    return o.transpose(1, 2).reshape(b, c, h, w)


def classifier(sd: dict, f: torch.Tensor, prefix="classifier.") -> torch.Tensor:
    for lin, ln in ((0, 1), (4, 5), (8, 9)):
        f = F.linear(f, sd[f"{prefix}{lin}.weight"], sd[f"{prefix}{lin}.bias"])
        f = F.relu(F.layer_norm(f, (f.shape[-1],), sd[f"{prefix}{ln}.weight"], sd[f"{prefix}{ln}.bias"], eps=1e-5))
    return F.linear(f, sd[prefix + "12.weight"], sd[prefix + "12.bias"])


def cross_attention(sd: dict, query: torch.Tensor, key_value: torch.Tensor, heads: int = 8,
                    prefix="cross_attention.") -> torch.Tensor:
    """CrossAttention.forward (modules.py:105-122): one query per image over the spatial tokens, + residual."""
    b = query.shape[0]
    emb = sd[prefix + "q_proj.weight"].shape[0]
    hd = emb // heads

    def proj(name, t):
        return F.linear(t, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"])

    q = proj("q_proj", query).unsqueeze(1).reshape(b, 1, heads, hd).transpose(1, 2)
    k = proj("k_proj", key_value).reshape(b, -1, heads, hd).transpose(1, 2)
    v = proj("v_proj", key_value).reshape(b, -1, heads, hd).transpose(1, 2)
    p = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(b, emb)
    return proj("out_proj", o) + query


def attention_decoder_logits(sd: dict, latent: torch.Tensor, heads: int = 8, use_spatial_attention=True,
                             use_self_attention=True, use_cross_attention=False) -> torch.Tensor:
    """AttentionClassificationDecoder.forward (modules.py:424-468)."""
    x = latent
    if use_spatial_attention:
        x = spatial_attention(sd, x)
    x = feature_compress(sd, x)
    if use_self_attention:
        x = self_attention(sd, x, heads)
    flat = x.reshape(x.shape[0], -1)
    if use_cross_attention:   # modules.py:450-459
        query = F.linear(flat, sd["query_generator.weight"], sd["query_generator.bias"])
        tokens = x.reshape(x.shape[0], x.shape[1], -1).transpose(1, 2)
        attended = cross_attention(sd, query, tokens, heads)
        flat = flat + attended.mean(dim=1, keepdim=True).expand_as(flat)
    return classifier(sd, flat)


def plain_decoder_logits(sd: dict, latent: torch.Tensor, use_adaptive_pooling=True) -> torch.Tensor:
    """ClassificationDecoder.forward (modules.py:335-349)."""
    f = (F.adaptive_avg_pool2d(latent, (4, 4)) if use_adaptive_pooling else latent).reshape(latent.shape[0], -1)
    for lin, ln in ((0, 1), (4, 5)):
        f = F.linear(f, sd[f"classifier.{lin}.weight"], sd[f"classifier.{lin}.bias"])
        f = F.layer_norm(f, (f.shape[-1],), sd[f"classifier.{ln}.weight"], sd[f"classifier.{ln}.bias"], eps=1e-5)
        f = F.leaky_relu(f, 0.2)
    return F.linear(f, sd["classifier.8.weight"], sd["classifier.8.bias"])


def get_confidence(logits: torch.Tensor):
    """modules.py:470-475: sigmoid then descending sort along tags."""
    conf = torch.sigmoid(logits)
    return torch.sort(conf, descending=True)


def threshold_tags(sorted_conf: torch.Tensor, indices: torch.Tensor, thr: float):
    """infer_full.py:109-124 for one image (1-D inputs): tags with conf >= thr, in sorted
    order, plus count / max / mean-of-top-5 with the reference's 4-decimal rounding."""
    conf = [float(c) for c in sorted_conf.tolist()]
    idx = [int(i) for i in indices.tolist()]
    picked = [(i, float(f"{c:.4f}")) for c, i in zip(conf, idx) if c >= thr]
    return {
        "predicted": picked,
        "total_tags_above_threshold": len(picked),
        "max_confidence": float(f"{max(conf):.4f}"),
        "avg_confidence_top5": float(f"{sum(conf[:5]) / 5:.4f}"),
    }


def focal_loss(logits: torch.Tensor, targets: torch.Tensor, alpha=1.0, gamma=2.0, reduction="mean"):
    """improved_losses.py:47-56."""
    bce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    pt = torch.exp(-bce)
    fl = alpha * (1.0 - pt) ** gamma * bce
    if reduction == "mean":
        return fl.mean()
    if reduction == "sum":
        return fl.sum()
    return fl


def class_balanced_loss(logits: torch.Tensor, targets: torch.Tensor, samples_per_class, beta=0.9999):
    """improved_losses.py:58-72 (ClassBalancedLoss.forward; gamma is unused there): effective-number class
    weights in float64 numpy, cast to float32, times the elementwise BCE, mean."""
    import numpy as np

    effective_num = 1.0 - np.power(beta, np.asarray(samples_per_class))
    weights = (1.0 - beta) / effective_num
    weights = weights / weights.sum() * len(weights)
    w = torch.tensor(weights, dtype=torch.float32)
    bce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    return (bce * w.to(bce.dtype).unsqueeze(0)).mean()


def focal_loss_grad(logits: torch.Tensor, targets: torch.Tensor, alpha=1.0, gamma=2.0):
    """d mean(focal)/d logits via autograd on the restatement above (for the fused kernel)."""
    x = logits.detach().clone().requires_grad_(True)
    focal_loss(x, targets, alpha, gamma).backward()
    return x.grad


# ----------------------------------------------------------------------------- train mode
CLASSIFIER_DROPOUT = (0.3, 0.2, 0.1)          # modules.py:405,410,415
PLAIN_CLASSIFIER_DROPOUT = (0.3, 0.2)         # modules.py:321,326


def _drop(x: torch.Tensor, mask, p: float) -> torch.Tensor:
    """nn.Dropout(p) in train mode with an explicit keep-mask (None = identity)."""
    return x if mask is None else x * mask / (1.0 - p)


def feature_compress_train(sd: dict, x: torch.Tensor, momentum=0.1, prefix="feature_compress."):
    """modules.py:377-382 in train mode.  Returns (pooled, new running_mean, new running_var)."""
    y = F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], padding=1)
    rm, rv = sd[prefix + "1.running_mean"].clone(), sd[prefix + "1.running_var"].clone()
    y = F.batch_norm(y, rm, rv, sd[prefix + "1.weight"], sd[prefix + "1.bias"], training=True, momentum=momentum,
                     eps=1e-5)
    return F.adaptive_avg_pool2d(F.relu(y), (8, 8)), rm, rv


def self_attention_train(sd: dict, x: torch.Tensor, heads: int = 8, attn_mask=None, p=0.1,
                         prefix="self_attention_post.") -> torch.Tensor:
    """modules.py:66-91 with the dropout of :81 applied through ``attn_mask`` [B,heads,N,N]."""
    b, c, h, w = x.shape
    n, hd = h * w, c // heads
    xf = x.reshape(b, c, n).transpose(1, 2)
    t = F.layer_norm(xf, (c,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps=1e-5)

    def proj(name):
        y = F.linear(t, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"])
        return y.reshape(b, n, heads, hd).transpose(1, 2)

    q, k, v = proj("q_proj"), proj("k_proj"), proj("v_proj")
    pw = _drop(torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd), dim=-1), attn_mask, p)
    o = torch.matmul(pw, v).transpose(1, 2).reshape(b, n, c)
    o = F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"]) + xf
    return o.transpose(1, 2).reshape(b, c, h, w)


def attention_decoder_train(sd: dict, latent: torch.Tensor, heads: int = 8, use_spatial_attention=True,
                            use_self_attention=True, attn_mask=None, attention_dropout=0.1, cls_masks=None,
                            momentum=0.1, use_cross_attention=False):
    """AttentionClassificationDecoder.forward (modules.py:424-468) under module.train().
    Returns (logits, new running_mean, new running_var)."""
    x = latent
    if use_spatial_attention:
        x = spatial_attention(sd, x)
    x, rm, rv = feature_compress_train(sd, x, momentum)
    if use_self_attention:
        x = self_attention_train(sd, x, heads, attn_mask, attention_dropout)
    f = x.reshape(x.shape[0], -1)
    if use_cross_attention:   # modules.py:450-459 (no dropout in this branch)
        query = F.linear(f, sd["query_generator.weight"], sd["query_generator.bias"])
        tokens = x.reshape(x.shape[0], x.shape[1], -1).transpose(1, 2)
        attended = cross_attention(sd, query, tokens, heads)
        f = f + attended.mean(dim=1, keepdim=True).expand_as(f)
    for i, (lin, ln) in enumerate(((0, 1), (4, 5), (8, 9))):
        f = F.linear(f, sd[f"classifier.{lin}.weight"], sd[f"classifier.{lin}.bias"])
        f = F.relu(F.layer_norm(f, (f.shape[-1],), sd[f"classifier.{ln}.weight"], sd[f"classifier.{ln}.bias"],
                                eps=1e-5))
        f = _drop(f, None if cls_masks is None else cls_masks[i], CLASSIFIER_DROPOUT[i])
    return F.linear(f, sd["classifier.12.weight"], sd["classifier.12.bias"]), rm, rv


def plain_decoder_train(sd: dict, latent: torch.Tensor, cls_masks=None, use_adaptive_pooling=True) -> torch.Tensor:
    """ClassificationDecoder.forward (modules.py:335-349) under module.train()."""
    f = (F.adaptive_avg_pool2d(latent, (4, 4)) if use_adaptive_pooling else latent).reshape(latent.shape[0], -1)
    for i, (lin, ln) in enumerate(((0, 1), (4, 5))):
        f = F.linear(f, sd[f"classifier.{lin}.weight"], sd[f"classifier.{lin}.bias"])
        f = F.layer_norm(f, (f.shape[-1],), sd[f"classifier.{ln}.weight"], sd[f"classifier.{ln}.bias"], eps=1e-5)
        f = _drop(F.leaky_relu(f, 0.2), None if cls_masks is None else cls_masks[i], PLAIN_CLASSIFIER_DROPOUT[i])
    return F.linear(f, sd["classifier.8.weight"], sd["classifier.8.bias"])


NON_TRAINABLE = ("running_mean", "running_var", "num_batches_tracked")


def head_train_step(sd: dict, latent, targets, alpha=1.0, gamma=2.0, kind="attention", dtype=torch.float32,
                    samples_per_class=None, **kw):
    """One reference training step on the head (train_decoder.py:186-195 without the optimizer):
    train-mode forward, FocalLoss(alpha, gamma) mean, autograd backward.
    Returns dict(loss, logits, grads{key: tensor}, running_mean, running_var).
    ``dtype=torch.float64`` evaluates the same graph in double precision (a tighter ground truth for
    gradients that are long cancelling sums); results are returned in that dtype."""
    def cast(v):
        return v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v

    latent, targets = cast(latent), cast(targets)
    kw = {k: ([cast(m) for m in v] if isinstance(v, (list, tuple)) else cast(v)) for k, v in kw.items()}
    leaf = {k: (cast(v.detach().clone()).requires_grad_(True) if not k.endswith(NON_TRAINABLE)
                else cast(v.detach().clone())) for k, v in sd.items()}
    if kind == "attention":
        logits, rm, rv = attention_decoder_train(leaf, latent, **kw)
    else:
        logits, rm, rv = plain_decoder_train(leaf, latent, **kw), None, None
    if samples_per_class is not None:     # train_decoder.py:188-189: ClassBalancedLoss instead of the focal loss
        loss = class_balanced_loss(logits, targets, samples_per_class)
    else:
        loss = focal_loss(logits, targets, alpha, gamma)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}
    return {"loss": loss.detach(), "logits": logits.detach(), "grads": grads, "running_mean": rm, "running_var": rv}


def adamw_step(p, g, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=1e-2, step=1, max_norm=0.0):
    """clip_grad_norm_(max_norm) then torch.optim.AdamW's update on flat tensors (train_decoder.py:197-203).
    Returns (p, m, v, grad norm before clipping)."""
    norm = g.double().norm().float()
    if max_norm > 0:
        g = g * torch.clamp(max_norm / (norm + 1e-6), max=1.0)
    p = p * (1.0 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    denom = v.sqrt() / math.sqrt(1 - beta2 ** step) + eps
    return p - (lr / (1 - beta1 ** step)) * m / denom, m, v, norm
