"""GPU-side image preprocessing + bucket batcher in front of the encoder.

The reference preprocesses every image on the host with PIL / torchvision (``infer_full.py:97-98``,
``modules.py:125-178``): ``SmartResize`` (centre crop to the bucket's aspect ratio + LANCZOS resize) or
``transforms.Resize((res, res))`` (PIL BILINEAR), then ``ToTensor`` and ``Normalize(0.5, 0.5)``.  Here the
decoded uint8 image is uploaded once (3 bytes per pixel instead of 12) and

  * crop + resize run as CUDA kernels that reproduce Pillow's 8-bit resampling bit for bit
    (``vt_resize_u8``), writing straight into the image's slot of a ``[B, H, W, 3]`` uint8 batch buffer;
  * ``ToTensor`` + ``Normalize`` are fused into the conv_in gather of ``vt_encode`` (``VT_IN_U8_NHWC``).

``BucketBatcher`` groups a stream of images by target shape (``AspectRatioBucketing.bucket_for_size`` or a
fixed square) and yields full device batches per bucket.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import _native
from .modules import AspectRatioBucketing


def gpu_smart_resize(img_u8: torch.Tensor, width: int, height: int, out: Optional[torch.Tensor] = None,
                     ctx: Optional[_native.Context] = None) -> torch.Tensor:
    """``SmartResize(width, height)(PIL image)`` for a uint8 [h,w,3] CUDA tensor -> uint8 [height,width,3]."""
    ctx = ctx or _native.get_context(img_u8.device)
    box = _native.smart_crop_box(img_u8.shape[1], img_u8.shape[0], width, height)
    return ctx.resize_u8(img_u8, (width, height), box, _native.FILTER_LANCZOS, out)


def gpu_square_resize(img_u8: torch.Tensor, resolution: int, out: Optional[torch.Tensor] = None,
                      ctx: Optional[_native.Context] = None) -> torch.Tensor:
    """``transforms.Resize((resolution, resolution))(PIL image)`` (PIL BILINEAR, aspect ratio not kept)."""
    ctx = ctx or _native.get_context(img_u8.device)
    return ctx.resize_u8(img_u8, (resolution, resolution), None, _native.FILTER_BILINEAR, out)


class BucketBatcher:
    """Collects decoded images (numpy / torch uint8 [h,w,3], host memory) and yields device batches.

    ``add(key, image)`` uploads the image, resizes it on the GPU into the next free slot of its bucket's
    batch buffer and returns a finished ``(shape, keys, batch_u8 [B,H,W,3])`` when the bucket is full;
    ``flush()`` yields the partial batches.  ``bucketing=None`` -> the fixed square ``resolution`` with
    the reference's BILINEAR ``Resize``; else SmartResize into the image's aspect-ratio bucket."""

    def __init__(self, device, batch_size: int = 8, resolution: int = 1024,
                 bucketing: Optional[AspectRatioBucketing] = None):
        self.device = torch.device(device)
        self.ctx = _native.get_context(self.device)
        self.batch_size, self.resolution, self.bucketing = batch_size, resolution, bucketing
        self._open: Dict[Tuple[int, int], Tuple[torch.Tensor, List]] = {}

    def shape_for(self, width: int, height: int) -> Tuple[int, int]:
        if self.bucketing is None:
            return (self.resolution, self.resolution)
        return self.bucketing.bucket_for_size(width, height)

    def add(self, key, image) -> Optional[Tuple[Tuple[int, int], List, torch.Tensor]]:
        img = torch.from_numpy(np.ascontiguousarray(image)) if isinstance(image, np.ndarray) else image.contiguous()
        if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
            raise ValueError("image must be uint8 [h, w, 3] (RGB)")
        h, w = img.shape[0], img.shape[1]
        shape = self.shape_for(w, h)
        W, H = shape
        if shape not in self._open:
            self._open[shape] = (torch.empty(self.batch_size, H, W, 3, dtype=torch.uint8, device=self.device), [])
        buf, keys = self._open[shape]
        dev = img.to(self.device, non_blocking=True)
        if self.bucketing is None:
            gpu_square_resize(dev, self.resolution, out=buf[len(keys)], ctx=self.ctx)
        else:
            gpu_smart_resize(dev, W, H, out=buf[len(keys)], ctx=self.ctx)
        keys.append(key)
        if len(keys) == self.batch_size:
            del self._open[shape]
            return shape, keys, buf
        return None

    def flush(self) -> Iterator[Tuple[Tuple[int, int], List, torch.Tensor]]:
        for shape, (buf, keys) in list(self._open.items()):
            if keys:
                yield shape, keys, buf[:len(keys)]
        self._open.clear()

    def batches(self, items: Iterable) -> Iterator[Tuple[Tuple[int, int], List, torch.Tensor]]:
        """items: iterable of (key, image).  Yields every full batch, then the partial ones."""
        for key, image in items:
            done = self.add(key, image)
            if done is not None:
                yield done
        yield from self.flush()
