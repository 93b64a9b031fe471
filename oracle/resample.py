"""Integer CPU restatement of the image preprocessing in front of the encoder.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference lines followed:
  * SmartResize.__call__ (crop box, then ``img.resize((W, H), Image.LANCZOS)``) ... modules.py:142-178
  * get_image_transform, square path ``transforms.Resize((res, res))`` -> PIL BILINEAR . modules.py:125-139
The arithmetic itself lives in a third-party dependency that is not vendored in the reference:
**Pillow** (``Pillow>=9.0`` unpinned in requirements.txt; this image has 12.2.0),
``src/libImaging/Resample.c``: ``precompute_coeffs`` (double-precision filter taps, normalised per output
pixel), ``normalize_coeffs_8bpc`` (taps rounded to fixed point with 22 fractional bits),
``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc`` (int32 accumulation started at
2^21, arithmetic shift by 22, clip to [0,255]; the intermediate image is uint8), the rule of
``ImagingResampleInner`` that a pass is skipped when that dimension does not change, and the rule of
``PIL/Image.py`` ``Image.resize`` (12.2) that an image more than 100 times taller than wide which shrinks
vertically gets its vertical pass first.

**Pinned**: ``tests/test_oracle_golden.py::test_resample_matches_pillow`` checks these functions bit for bit
against Pillow itself (``Image.crop(...).resize(...)``) on seeded random images, in this container and
wherever the tests run (Pillow ships in the image; it is a library, not the reference).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS, BILINEAR = 1, 2          # PIL.Image.Resampling values
_SUPPORT = {LANCZOS: 3.0, BILINEAR: 1.0}


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _filter(kind: int, x: float) -> float:
    if kind == LANCZOS:
        return _sinc(x) * _sinc(x / 3) if -3.0 <= x < 3.0 else 0.0
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size: int, out_size: int, kind: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size).
    Returns (ksize, bounds[out,2] int32 (first tap, tap count), kk[out,ksize] int32)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = _SUPPORT[kind] * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_filter(kind, (x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, out_size: int, kind: int, axis: int) -> np.ndarray:
    """One 8-bit pass along ``axis`` (1 = horizontal, 0 = vertical) of an [h, w, c] uint8 image."""
    in_size = img.shape[axis]
    _, bounds, kk = precompute_coeffs(in_size, out_size, kind)
    src = np.moveaxis(img, axis, 0).astype(np.int64)          # [in, other, c]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        acc = ((acc + 2 ** 31) % 2 ** 32 - 2 ** 31)           # int32 wrap-around like the C accumulator
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img: np.ndarray, out_w: int, out_h: int, kind: int = LANCZOS) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((out_w, out_h), kind)`` for an [h, w, c] uint8 array."""
    h, w = img.shape[:2]
    if h > w * 100 and out_h < h:
        # PIL/Image.py (Pillow 12.2) Image.resize: a tall thin image that shrinks vertically is resized in two
        # calls, vertical pass first
        img = _pass(img, out_h, kind, 0)
        if w != out_w:
            img = _pass(img, out_w, kind, 1)
        return np.ascontiguousarray(img)
    if w != out_w:
        img = _pass(img, out_w, kind, 1)
    if h != out_h:
        img = _pass(img, out_h, kind, 0)
    return np.ascontiguousarray(img)


def smart_crop_box(ow: int, oh: int, tw: int, th: int, crop_mode: str = "center"):
    """Crop box (left, top, right, bottom) of SmartResize (modules.py:149-172); 'random' is not restated."""
    target = tw / th
    ratio = ow / oh
    if ratio > target:
        nw = int(oh * target)
        left = (ow - nw) // 2 if crop_mode == "center" else 0
        return left, 0, left + nw, oh
    if ratio < target:
        nh = int(ow / target)
        top = (oh - nh) // 2 if crop_mode == "center" else 0
        return 0, top, ow, top + nh
    return 0, 0, ow, oh


def smart_resize_u8(img: np.ndarray, tw: int, th: int, crop_mode: str = "center") -> np.ndarray:
    """SmartResize(tw, th, crop_mode)(PIL image) as arrays: crop to the target ratio, LANCZOS resize."""
    l, t, r, b = smart_crop_box(img.shape[1], img.shape[0], tw, th, crop_mode)
    return resize_u8(img[t:b, l:r], tw, th, LANCZOS)
