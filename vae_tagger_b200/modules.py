"""Drop-in for the tag-decoder half of the reference's ``modules.py``.

Public names, constructor arguments and state-dict keys follow ``/root/reference/modules.py``
(``SpatialAttention`` :15-47, ``MultiHeadSelfAttention`` :49-91, ``CrossAttention`` :93-124,
``ClassificationDecoder`` :303-356, ``AttentionClassificationDecoder`` :358-485,
``create_attention_decoder`` :731-748 and the small helpers :126-301).

Inference (``eval()`` / ``get_confidence``) runs the hand-written CUDA head kernels through
the C-ABI -- one fused pipeline per batch plus an on-device sort -- and raises if the native
library or a CUDA device is missing.  ``train()`` mode keeps a differentiable PyTorch graph
for ``train_decoder.py`` (dropout, BatchNorm batch statistics, autograd): the head's backward
pass is not part of the inference hot path.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .autoencoder_kl import AutoencoderKL

IMAGE_EXTENSIONS = (".png", ".jpg", ".jpeg", ".bmp", ".tiff", ".webp")


# ----------------------------------------------------------------------------- attention blocks
class SpatialAttention(nn.Module):
    """Channel gate (shared 1x1 MLP over global avg / max pools) then spatial gate (7x7 conv over
    the channel mean / max maps).  Parameters: ``channel_att.{0,2}.weight``, ``spatial_att.0.weight``."""

    def __init__(self, in_channels, reduction_ratio=8):
        super().__init__()
        hidden = in_channels // reduction_ratio
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.channel_att = nn.Sequential(
            nn.Conv2d(in_channels, hidden, 1, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(hidden, in_channels, 1, bias=False))
        self.spatial_att = nn.Sequential(nn.Conv2d(2, 1, kernel_size=7, padding=3, bias=False), nn.Sigmoid())
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        gate_c = self.sigmoid(self.channel_att(self.avg_pool(x)) + self.channel_att(self.max_pool(x)))
        x = x * gate_c
        pooled = torch.cat([x.mean(dim=1, keepdim=True), x.amax(dim=1, keepdim=True)], dim=1)
        return x * self.spatial_att(pooled)


class MultiHeadSelfAttention(nn.Module):
    """Pre-LN self-attention over the h*w positions of a [B,C,h,w] map with a residual connection."""

    def __init__(self, embed_dim, num_heads=8, dropout=0.1):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.dropout = nn.Dropout(dropout)
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x):
        b, c, h, w = x.shape
        tokens = x.flatten(2).transpose(1, 2)
        t = self.norm(tokens)

        def split(p):
            return p(t).view(b, h * w, self.num_heads, self.head_dim).transpose(1, 2)

        q, k, v = split(self.q_proj), split(self.k_proj), split(self.v_proj)
        att = self.dropout(F.softmax(q @ k.transpose(-2, -1) / math.sqrt(self.head_dim), dim=-1))
        o = (att @ v).transpose(1, 2).reshape(b, h * w, c)
        return (self.out_proj(o) + tokens).transpose(1, 2).reshape(b, c, h, w)


class CrossAttention(nn.Module):
    """Single-query cross attention (optional ``--use_cross_attention`` branch, off by default)."""

    def __init__(self, query_dim, key_dim, embed_dim, num_heads=8):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.q_proj = nn.Linear(query_dim, embed_dim)
        self.k_proj = nn.Linear(key_dim, embed_dim)
        self.v_proj = nn.Linear(key_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, query_dim)

    def forward(self, query, key_value):
        b = query.shape[0]

        def heads(t):
            return t.view(b, -1, self.num_heads, self.head_dim).transpose(1, 2)

        q = heads(self.q_proj(query).unsqueeze(1))
        k, v = heads(self.k_proj(key_value)), heads(self.v_proj(key_value))
        att = F.softmax(q @ k.transpose(-2, -1) / math.sqrt(self.head_dim), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(b, self.embed_dim)
        return self.out_proj(o) + query


# ----------------------------------------------------------------------------- native dispatch
class _NativeHeadMixin:
    """Mirrors the module's parameters into the native context and runs the CUDA head."""

    _native_key = None

    def _head_config(self):  # -> (kind, kwargs)
        raise NotImplementedError

    def _native_ctx(self, device):
        if device.type != "cuda":
            raise _native.NativeError(
                f"{type(self).__name__} inference needs CUDA tensors on a B200: the tag head has no CPU fallback")
        ctx = _native.get_context(device)
        sd = self.state_dict()
        key = (id(ctx), tuple((k, v.data_ptr(), v._version) for k, v in sd.items()))
        if key != self._native_key or getattr(ctx, "_head_owner", None) is not self:
            kind, kw = self._head_config()
            ctx.configure_head(kind, **kw)
            ctx.load_head(sd)
            ctx._head_owner = self
            self._native_key = key
        return ctx

    def _use_native(self, x=None) -> bool:
        """Eval mode runs the CUDA head.  ``train()`` mode keeps the module's own graph (batch-statistics BatchNorm
        and Dropout, as the reference -- also under ``no_grad``, e.g. ``get_confidence`` on a training-mode
        decoder), and so does an eval-mode call whose input needs a gradient (the kernels build no autograd graph).
        The native TRAINING step does not come through here: ``DecoderTrainer`` calls ``vt_head_train_step``."""
        if self.training:
            return False
        return not (x is not None and torch.is_grad_enabled() and x.requires_grad)

    def get_confidence(self, latent_vectors):
        """sigmoid(logits) sorted descending with the tag indices (reference ``get_confidence``)."""
        if self._use_native(latent_vectors):
            out = self._native_ctx(latent_vectors.device).tag(latent_vectors, want=("conf", "idx"))
            return out["conf"], out["idx"]
        with torch.no_grad():
            conf = torch.sigmoid(self(latent_vectors))
            return torch.sort(conf, descending=True)

    @torch.no_grad()
    def tag(self, latent_vectors, threshold=0.5):
        """One call for what ``infer_full.py:102-118`` does per image: returns a dict with the sorted
        confidences, indices and the per-image count of ``conf >= threshold`` -- all on the device."""
        return self._native_ctx(latent_vectors.device).tag(latent_vectors, threshold=threshold,
                                                           want=("conf", "idx", "count"))


class ClassificationDecoder(_NativeHeadMixin, nn.Module):
    """Plain head (``--no_attention``): AdaptiveAvgPool(4,4) -> 256-512-256-T MLP with LayerNorm /
    LeakyReLU(0.2) / Dropout."""

    def __init__(self, latent_channels, latent_height, latent_width, num_classes, use_adaptive_pooling=True):
        super().__init__()
        self.latent_channels, self.latent_height, self.latent_width = latent_channels, latent_height, latent_width
        self.num_classes = num_classes
        self.use_adaptive_pooling = use_adaptive_pooling
        if use_adaptive_pooling:
            self.adaptive_pool = nn.AdaptiveAvgPool2d((4, 4))
            in_dim = latent_channels * 16
        else:
            in_dim = latent_channels * latent_height * latent_width
        self.classifier = nn.Sequential(
            nn.Linear(in_dim, 512), nn.LayerNorm(512), nn.LeakyReLU(0.2), nn.Dropout(0.3),
            nn.Linear(512, 256), nn.LayerNorm(256), nn.LeakyReLU(0.2), nn.Dropout(0.2),
            nn.Linear(256, num_classes))

    def _head_config(self):
        flat = 0 if self.use_adaptive_pooling else self.latent_channels * self.latent_height * self.latent_width
        return _native.HEAD_PLAIN, dict(latent_channels=self.latent_channels, num_classes=self.num_classes,
                                        plain_flat_dim=flat)

    def forward(self, latent_vectors):
        if self._use_native(latent_vectors):
            return self._native_ctx(latent_vectors.device).tag(latent_vectors, want=("logits",))["logits"]
        x = self.adaptive_pool(latent_vectors) if self.use_adaptive_pooling else latent_vectors
        return self.classifier(x.reshape(latent_vectors.size(0), -1))


class AttentionClassificationDecoder(_NativeHeadMixin, nn.Module):
    """SpatialAttention -> conv3x3/BN/ReLU/AdaptiveAvgPool(8,8) -> self-attention on the 8x8 grid ->
    512-1024-512-256-T MLP (LayerNorm/ReLU/Dropout) -> per-tag logits."""

    def __init__(self, latent_channels, latent_height, latent_width, num_classes, use_spatial_attention=True,
                 use_self_attention=True, use_cross_attention=False, attention_heads=8, attention_dropout=0.1):
        super().__init__()
        self.latent_channels, self.latent_height, self.latent_width = latent_channels, latent_height, latent_width
        self.num_classes = num_classes
        self.use_spatial_attention = use_spatial_attention
        self.use_self_attention = use_self_attention
        self.use_cross_attention = use_cross_attention
        self.attention_heads = attention_heads
        half = latent_channels // 2
        if use_spatial_attention:
            self.spatial_attention = SpatialAttention(latent_channels)
        self.feature_compress = nn.Sequential(
            nn.Conv2d(latent_channels, half, 3, 1, 1), nn.BatchNorm2d(half), nn.ReLU(inplace=True),
            nn.AdaptiveAvgPool2d((8, 8)))
        flat = half * 64
        if use_self_attention:
            self.self_attention_post = MultiHeadSelfAttention(half, num_heads=attention_heads,
                                                              dropout=attention_dropout)
        if use_cross_attention:
            self.cross_attention = CrossAttention(query_dim=512, key_dim=half, embed_dim=256,
                                                  num_heads=attention_heads)
        self.classifier = nn.Sequential(
            nn.Linear(flat, 1024), nn.LayerNorm(1024), nn.ReLU(inplace=True), nn.Dropout(0.3),
            nn.Linear(1024, 512), nn.LayerNorm(512), nn.ReLU(inplace=True), nn.Dropout(0.2),
            nn.Linear(512, 256), nn.LayerNorm(256), nn.ReLU(inplace=True), nn.Dropout(0.1),
            nn.Linear(256, num_classes))
        if use_cross_attention:
            self.query_generator = nn.Linear(flat, 512)

    def _head_config(self):
        return _native.HEAD_ATTENTION, dict(
            latent_channels=self.latent_channels, num_classes=self.num_classes,
            use_spatial_attention=self.use_spatial_attention, use_self_attention=self.use_self_attention,
            attention_heads=self.attention_heads, use_cross_attention=self.use_cross_attention)

    def forward(self, latent_vectors):
        if self._use_native(latent_vectors):
            return self._native_ctx(latent_vectors.device).tag(latent_vectors, want=("logits",))["logits"]
        x = latent_vectors
        if self.use_spatial_attention:
            x = self.spatial_attention(x)
        x = self.feature_compress(x)
        if self.use_self_attention:
            x = self.self_attention_post(x)
        flat = x.reshape(x.size(0), -1)
        if self.use_cross_attention:
            attended = self.cross_attention(self.query_generator(flat), x.flatten(2).transpose(1, 2))
            flat = flat + attended.mean(dim=1, keepdim=True).expand_as(flat)
        return self.classifier(flat)

    def get_attention_maps(self, latent_vectors):
        return {}


def create_attention_decoder(latent_channels, latent_height, latent_width, num_classes, attention_config=None):
    """``attention_config=None`` -> plain head, else the attention head with the reference's defaults."""
    if attention_config is None:
        return ClassificationDecoder(latent_channels, latent_height, latent_width, num_classes)
    cfg = attention_config
    return AttentionClassificationDecoder(
        latent_channels=latent_channels, latent_height=latent_height, latent_width=latent_width,
        num_classes=num_classes, use_spatial_attention=cfg.get("use_spatial_attention", True),
        use_self_attention=cfg.get("use_self_attention", True),
        use_cross_attention=cfg.get("use_cross_attention", False),
        attention_heads=cfg.get("attention_heads", 8), attention_dropout=cfg.get("attention_dropout", 0.1))


# ----------------------------------------------------------------------------- helpers (host side)
def get_vae_latent_info(resolution, latent_channels=16):
    side = resolution // 8
    return {"latent_channels": latent_channels, "latent_height": side, "latent_width": side,
            "total_dim": latent_channels * side * side}


def get_vae_config(resolution, use_quant_conv, use_post_quant_conv):
    return {
        "in_channels": 3, "out_channels": 3, "down_block_types": ["DownEncoderBlock2D"] * 4,
        "up_block_types": ["UpDecoderBlock2D"] * 4, "block_out_channels": [128, 256, 512, 512],
        "layers_per_block": 2, "act_fn": "silu", "latent_channels": 16, "norm_num_groups": 32,
        "sample_size": resolution, "mid_block_add_attention": True, "use_quant_conv": use_quant_conv,
        "use_post_quant_conv": use_post_quant_conv,
    }


def get_image_paths(path):
    """Image files under a directory (recursive, de-duplicated) or a single image file.  The reference returns
    ``list(set(...))`` (modules.py:270) -- an arbitrary, per-process order; here the list is sorted, so that every
    rank of a sharded run sees the same order and the output is reproducible."""
    if os.path.isdir(path):
        # the reference globs "*<ext>" and "*<EXT>" (modules.py:263-268): all-lower or all-upper extensions only
        both = IMAGE_EXTENSIONS + tuple(e.upper() for e in IMAGE_EXTENSIONS)
        found = {p.resolve() for p in Path(path).rglob("*") if p.is_file() and p.name.endswith(both)}
        return sorted(found)
    if os.path.isfile(path):
        if path.lower().endswith(IMAGE_EXTENSIONS):
            return [Path(path)]
        print(f"warning: {path} is not a supported image format")
        return []
    print(f"error: path {path} does not exist")
    return []


class AspectRatioBucketing:
    """(width, height) buckets from ``base_resolution`` to ``max_resolution`` in ``bucket_step`` steps
    whose area does not exceed ``max_resolution**2``; an image goes to the bucket with the nearest
    aspect ratio, the first one in sorted order winning ties."""

    def __init__(self, base_resolution=512, max_resolution=1024, bucket_step=64):
        self.base_resolution, self.max_resolution, self.bucket_step = base_resolution, max_resolution, bucket_step
        self.buckets = self._generate_buckets()
        self.image_buckets = {}

    def _generate_buckets(self):
        sides = range(self.base_resolution, self.max_resolution + 1, self.bucket_step)
        area = self.max_resolution * self.max_resolution
        return sorted((w, h) for w in sides for h in sides if w * h <= area)

    def bucket_for_size(self, width, height):
        ratio = width / height
        best, best_diff = None, float("inf")
        for w, h in self.buckets:
            d = abs(w / h - ratio)
            if d < best_diff:
                best, best_diff = (w, h), d
        return best

    def assign_bucket(self, image_path):
        try:
            from PIL import Image

            with Image.open(image_path) as img:
                bucket = self.bucket_for_size(*img.size)
        except Exception as e:  # noqa: BLE001
            print(f"warning: cannot analyse image {image_path}: {e}")
            return (self.base_resolution, self.base_resolution)
        self.image_buckets[image_path] = bucket
        return bucket

    def get_bucket_statistics(self):
        counts = {}
        for b in self.image_buckets.values():
            counts[b] = counts.get(b, 0) + 1
        return counts

    def print_bucket_info(self):
        stats = self.get_bucket_statistics()
        print(f"aspect-ratio buckets: {len(self.buckets)} generated, {len(stats)} in use")
        for (w, h), n in sorted(stats.items(), key=lambda kv: -kv[1]):
            print(f"  {w}x{h} (ratio {w / h:.2f}): {n} images ({100.0 * n / max(1, len(self.image_buckets)):.1f}%)")


class SmartResize:
    """Crop to the target aspect ratio (centre / random / top-left), then LANCZOS-resize."""

    def __init__(self, target_width, target_height, crop_mode="center"):
        self.target_width, self.target_height, self.crop_mode = target_width, target_height, crop_mode

    def _offset(self, slack):
        if self.crop_mode == "center":
            return slack // 2
        if self.crop_mode == "random":
            import random

            return random.randint(0, slack)
        return 0

    def __call__(self, img):
        from PIL import Image

        ow, oh = img.size
        target = self.target_width / self.target_height
        ratio = ow / oh
        if ratio > target:
            nw = int(oh * target)
            left = self._offset(ow - nw)
            img = img.crop((left, 0, left + nw, oh))
        elif ratio < target:
            nh = int(ow / target)
            top = self._offset(oh - nh)
            img = img.crop((0, top, ow, top + nh))
        return img.resize((self.target_width, self.target_height), Image.LANCZOS)


def get_image_transform(resolution, use_bucketing=False, aspect_ratio_bucket=None):
    from torchvision import transforms

    if use_bucketing and aspect_ratio_bucket is not None:
        w, h = aspect_ratio_bucket
        first = SmartResize(w, h)
    else:
        first = transforms.Resize((resolution, resolution))
    return transforms.Compose([first, transforms.ToTensor(), transforms.Normalize([0.5] * 3, [0.5] * 3)])


class VAE(nn.Module):
    """``VAE(vae_config).encode(x)`` = ``latent_dist.mode()`` (reference modules.py:288-301)."""

    def __init__(self, vae_config):
        super().__init__()
        self.vae = AutoencoderKL(**vae_config)

    def forward(self, x):
        posterior = self.vae.encode(x).latent_dist
        return self.vae.decode(posterior.sample()).sample, posterior

    def encode(self, x):
        return self.vae.encode(x).latent_dist.mode()


class TaggedImageDataset(torch.utils.data.Dataset):
    """``{image_path: "tag[:w], tag[:w], ..."}`` JSON + ``tags.csv`` (column ``name``) -> dicts with
    ``pixel_values`` and multi-hot ``labels`` (the two keys ``train_decoder.py`` consumes).  With
    ``triplets=True`` every item also carries ``anchor`` / ``positive`` / ``negative`` images and
    ``positive_labels`` / ``negative_labels`` from the reference's online triplet mining (modules.py:600-685),
    which ``train_full.py`` / ``train_vae.py`` consume."""

    def __init__(self, json_path, tags_csv_path, transform=None, use_bucketing=False, base_resolution=512,
                 max_resolution=1024, bucket_step=64, raw_uint8=False, triplets=False):
        import json

        self.triplets = triplets

        # raw_uint8 (addition of this implementation): ``pixel_values`` is the decoded image as a uint8 [h,w,3]
        # tensor plus its target (W, H) under ``target_size``; resize / crop / normalise then run on the GPU
        # (vae_tagger_b200.preprocess), bit-exact with the PIL transform, instead of in the loader workers
        self.raw_uint8 = raw_uint8

        import pandas as pd

        with open(json_path, "r") as f:
            self.data = json.load(f)
        self.tags = list(pd.read_csv(tags_csv_path)["name"])
        self.tag_to_idx = {t: i for i, t in enumerate(self.tags)}
        self.idx_to_tag = {i: t for t, i in self.tag_to_idx.items()}
        self.transform = transform
        self.image_paths = list(self.data.keys())
        self.use_bucketing = use_bucketing
        self.bucketing = AspectRatioBucketing(base_resolution, max_resolution, bucket_step) if use_bucketing else None
        if self.bucketing:
            for p in self.image_paths:
                self.bucketing.assign_bucket(p)
        self._bucket_tf = {}
        self.image_labels = {p: self._parse(prompt) for p, prompt in self.data.items()}

    def _parse(self, prompt):
        labels = torch.zeros(len(self.tags), dtype=torch.float32)
        for entry in prompt.split(","):
            name, _, weight = entry.partition(":")
            name = name.strip()
            try:
                w = float(weight.strip()) if weight.strip() else 1.0
            except ValueError:
                w = 1.0
            if name in self.tag_to_idx:
                labels[self.tag_to_idx[name]] = w
        return labels

    def __len__(self):
        return len(self.image_paths)

    def _load(self, path):
        from PIL import Image

        try:
            img = Image.open(path).convert("RGB")
        except Exception as e:  # noqa: BLE001
            print(f"warning: cannot load image {path}: {e}")
            side = 512 if self.use_bucketing else 224
            img = Image.new("RGB", (side, side), (0, 0, 0))
        bucket = self.bucketing.image_buckets.get(path) if self.bucketing else None
        if self.raw_uint8:
            import numpy as np

            return torch.from_numpy(np.array(img)), bucket   # a writable copy of the decoded pixels
        if bucket:
            if bucket not in self._bucket_tf:
                self._bucket_tf[bucket] = get_image_transform(0, True, bucket)
            return self._bucket_tf[bucket](img)
        if self.transform:
            return self.transform(img)
        return get_image_transform(512)(img) if self.use_bucketing else img

    def _mine_triplet(self, idx, anchor_labels, max_candidates=100):
        """Online triplet mining as the reference does it (modules.py:600-685): up to ``max_candidates`` random other
        images split into positives (any shared tag) and negatives; a multi-tag anchor takes the positive with the
        largest overlap 70 % of the time, otherwise a random one; without candidates the anchor stands in for the
        positive and a random other image for the negative."""
        import random

        n = len(self.image_paths)
        anchor = self.image_paths[idx]
        k = min(max_candidates, max(0, n - 1))
        picked = set()
        while len(picked) < k:
            j = random.randrange(0, n)
            if j != idx:
                picked.add(j)
        pos, neg = [], []
        for j in picked:
            p = self.image_paths[j]
            (pos if (self.image_labels[p] * anchor_labels).sum().item() > 0 else neg).append(p)
        if pos:
            if anchor_labels.sum().item() > 1 and len(pos) > 1 and random.random() < 0.7:
                positive = max(pos, key=lambda p: (self.image_labels[p] * anchor_labels).sum().item())
            else:
                positive = random.choice(pos)
        else:
            positive = anchor
        if neg:
            negative = random.choice(neg)
        elif n > 1:
            j = idx
            while j == idx:
                j = random.randrange(0, n)
            negative = self.image_paths[j]
        else:
            negative = anchor
        return positive, negative

    def __getitem__(self, idx):
        path = self.image_paths[idx]
        if self.raw_uint8:
            img, bucket = self._load(path)
            return {"pixel_values": img, "target_size": bucket, "labels": self.image_labels[path]}
        img = self._load(path)
        item = {"pixel_values": img, "labels": self.image_labels[path]}
        if self.triplets:
            labels = self.image_labels[path]
            pos, neg = self._mine_triplet(idx, labels)
            item.update(anchor=img, positive=self._load(pos), negative=self._load(neg),
                        positive_labels=self.image_labels.get(pos, labels),
                        negative_labels=self.image_labels.get(neg, torch.zeros_like(labels)))
        return item
