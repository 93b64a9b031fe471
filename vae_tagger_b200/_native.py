"""ctypes binding of ``include/vae_tagger_b200.h`` (the C-ABI of the encode+tag path).

There is NO fallback: if ``_lib/libvt_b200.so`` is missing or a call fails, a
``NativeError`` is raised.  Tensors cross the boundary as raw device pointers
(``tensor.data_ptr()``) plus sizes; the stream is torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _build

PREC_BF16, PREC_FP32, PREC_F16 = 0, 1, 2
IN_F32_NCHW, IN_U8_NHWC = 0, 1
HEAD_ATTENTION, HEAD_PLAIN = 0, 1
FILTER_LANCZOS, FILTER_BILINEAR = 1, 2   # PIL.Image.LANCZOS / BILINEAR
NUM_KERNEL_CLASSES = 13
KERNEL_CLASS_NAMES = ["igemm_tcgen05", "group_norm", "conv_in_gather", "softmax", "latent", "head", "fp32_contract", "misc",
                      "conv3_fused_128t", "conv3_fused_256", "flash_d512", "conv_in", "backward_contract"]
TENSOR_KERNEL_CLASSES = ("igemm_tcgen05", "conv3_fused_128t", "conv3_fused_256", "flash_d512", "conv_in")


class NativeError(RuntimeError):
    pass


class EncoderConfig(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int), ("num_blocks", C.c_int), ("block_out_channels", C.c_int * 8),
        ("layers_per_block", C.c_int), ("norm_num_groups", C.c_int), ("latent_channels", C.c_int),
        ("mid_block_add_attention", C.c_int), ("has_scaling_factor", C.c_int), ("has_shift_factor", C.c_int),
        ("scaling_factor", C.c_float), ("shift_factor", C.c_float),
    ]


class EncodeArgs(C.Structure):
    _fields_ = [
        ("images", C.c_void_p), ("in_fmt", C.c_int), ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("precision", C.c_int), ("sample", C.c_int), ("apply_scale_shift", C.c_int), ("seed", C.c_uint64),
        ("noise", C.c_void_p), ("latent", C.c_void_p), ("mean", C.c_void_p), ("logvar", C.c_void_p),
        ("micro_batch", C.c_int), ("single_lane", C.c_int), ("stream", C.c_void_p),
    ]


class DecodeArgs(C.Structure):
    _fields_ = [
        ("latent", C.c_void_p), ("batch", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int), ("precision", C.c_int),
        ("apply_scale_shift", C.c_int), ("image", C.c_void_p), ("micro_batch", C.c_int), ("stream", C.c_void_p),
    ]


class HeadConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("latent_channels", C.c_int), ("num_classes", C.c_int),
        ("use_spatial_attention", C.c_int), ("use_self_attention", C.c_int), ("attention_heads", C.c_int),
        ("use_cross_attention", C.c_int), ("plain_flat_dim", C.c_int),
    ]


class TagArgs(C.Structure):
    _fields_ = [
        ("latent", C.c_void_p), ("batch", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int), ("threshold", C.c_float),
        ("logits", C.c_void_p), ("probs", C.c_void_p), ("conf_sorted", C.c_void_p), ("idx_sorted", C.c_void_p),
        ("count", C.c_void_p), ("stream", C.c_void_p),
    ]


class InferHostArgs(C.Structure):
    _fields_ = [
        ("images_host", C.c_void_p), ("in_fmt", C.c_int), ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("precision", C.c_int), ("threshold", C.c_float), ("conf_sorted_host", C.c_void_p),
        ("idx_sorted_host", C.c_void_p), ("count_host", C.c_void_p), ("latent_host", C.c_void_p),
        ("micro_batch", C.c_int), ("stream", C.c_void_p),
    ]


class InferArgs(C.Structure):
    _fields_ = [
        ("images", C.c_void_p), ("in_fmt", C.c_int), ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("precision", C.c_int), ("threshold", C.c_float), ("conf_sorted", C.c_void_p), ("idx_sorted", C.c_void_p),
        ("count", C.c_void_p), ("latent", C.c_void_p), ("micro_batch", C.c_int), ("single_lane", C.c_int),
        ("stream", C.c_void_p),
    ]


class HeadTrainArgs(C.Structure):
    _fields_ = [
        ("latent", C.c_void_p), ("targets", C.c_void_p), ("batch", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int),
        ("params", C.c_void_p), ("grads", C.c_void_p), ("bn_running_mean", C.c_void_p),
        ("bn_running_var", C.c_void_p), ("bn_num_batches_tracked", C.c_void_p), ("bn_momentum", C.c_float),
        ("focal_alpha", C.c_float), ("focal_gamma", C.c_float), ("loss_scale", C.c_float), ("dropout", C.c_int),
        ("attention_dropout", C.c_float), ("seed", C.c_uint64), ("loss", C.c_void_p), ("logits", C.c_void_p),
        ("stream", C.c_void_p), ("class_weights", C.c_void_p),
    ]


class ResizeArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_w", C.c_int), ("src_h", C.c_int), ("src_stride", C.c_int64),
        ("crop_l", C.c_int), ("crop_t", C.c_int), ("crop_r", C.c_int), ("crop_b", C.c_int),
        ("dst", C.c_void_p), ("dst_w", C.c_int), ("dst_h", C.c_int), ("dst_stride", C.c_int64),
        ("filter", C.c_int), ("stream", C.c_void_p),
    ]


class EncoderBackwardArgs(C.Structure):
    _fields_ = [("grad_mean", C.c_void_p), ("grad_logvar", C.c_void_p), ("slot", C.c_int), ("accumulate", C.c_int),
                ("stream", C.c_void_p)]


class DecoderBackwardArgs(C.Structure):
    _fields_ = [("grad_image", C.c_void_p), ("grad_latent", C.c_void_p), ("slot", C.c_int), ("accumulate", C.c_int),
                ("stream", C.c_void_p)]


MAX_TAPES = 8


class EmbedLossArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("p", C.c_void_p), ("n", C.c_void_p), ("labels_a", C.c_void_p), ("labels_p", C.c_void_p),
        ("B", C.c_int), ("D", C.c_int64), ("T", C.c_int), ("kind", C.c_int), ("similarity", C.c_int),
        ("margin", C.c_float), ("loss", C.c_void_p), ("grad_a", C.c_void_p), ("grad_p", C.c_void_p),
        ("grad_n", C.c_void_p), ("stream", C.c_void_p),
    ]


class ResnetBlockPtrs(C.Structure):
    """vt_resnet_block_params (const float*) and, with one more leading field, vt_resnet_block_grads."""
    _fields_ = [(n, C.c_void_p) for n in ("norm1_w", "norm1_b", "conv1_w", "conv1_b", "norm2_w", "norm2_b", "conv2_w",
                                          "conv2_b", "sc_w", "sc_b")]


class ResnetBlockGrads(C.Structure):
    _fields_ = [("x", C.c_void_p)] + ResnetBlockPtrs._fields_


# every symbol include/vae_tagger_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "vt_last_error": (C.c_char_p, []),
    "vt_abi_version": (C.c_int, []),
    "vt_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vt_ctx_destroy": (C.c_int, [_P]),
    "vt_encoder_configure": (C.c_int, [_P, C.POINTER(EncoderConfig)]),
    "vt_encoder_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "vt_encoder_finalize": (C.c_int, [_P]),
    "vt_encode": (C.c_int, [_P, C.POINTER(EncodeArgs)]),
    "vt_decoder_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "vt_decoder_finalize": (C.c_int, [_P]),
    "vt_decode": (C.c_int, [_P, C.POINTER(DecodeArgs)]),
    "vt_head_configure": (C.c_int, [_P, C.POINTER(HeadConfig)]),
    "vt_head_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "vt_head_finalize": (C.c_int, [_P]),
    "vt_tag": (C.c_int, [_P, C.POINTER(TagArgs)]),
    "vt_infer_host": (C.c_int, [_P, C.POINTER(InferHostArgs)]),
    "vt_infer": (C.c_int, [_P, C.POINTER(InferArgs)]),
    "vt_focal_loss": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, _P, _P, _P]),
    "vt_head_param_count": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "vt_head_param_layout": (C.c_int, [_P, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]),
    "vt_head_train_step": (C.c_int, [_P, C.POINTER(HeadTrainArgs)]),
    "vt_head_dropout_masks": (C.c_int, [_P, C.c_int, C.c_float, C.c_uint64, _P, _P, _P, _P, _P]),
    "vt_adamw_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64] + [C.c_float] * 5 + [C.c_int64, C.c_float, C.c_float,
                                                                                  C.c_int, _P, _P]),
    "vt_resize_u8": (C.c_int, [_P, C.POINTER(ResizeArgs)]),
    "vt_resize_u8_batch": (C.c_int, [_P, C.POINTER(ResizeArgs), C.c_int, _P]),
    "vt_smart_crop_box": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "vt_resize_coefficients": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32)]),
    "vt_profile_enable": (C.c_int, [_P, C.c_int]),
    "vt_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "vt_op_conv2d": (C.c_int, [_P, _P, _P, _P, _P, _P, _P] + [C.c_int] * 9 + [_P, _P, _P]),
    "vt_op_conv3_fused": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P] + [C.c_int] * 6 + [C.c_float, C.c_int, _P, _P, _P]),
    "vt_op_flash_attention": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P, _P]),
    "vt_op_gemm_nt": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 5 + [C.c_float, C.c_int, _P, _P]),
    "vt_op_group_norm": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 5 + [C.c_float, C.c_int, C.c_int, _P, _P]),
    "vt_op_softmax_rows": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "vt_encoder_train_forward": (C.c_int, [_P, C.POINTER(EncodeArgs), C.c_int]),
    "vt_encoder_grad_bind": (C.c_int, [_P, C.c_char_p, _P]),
    "vt_encoder_backward": (C.c_int, [_P, C.POINTER(EncoderBackwardArgs)]),
    "vt_encoder_tape_release": (C.c_int, [_P, C.c_int]),
    "vt_decoder_train_forward": (C.c_int, [_P, C.POINTER(DecodeArgs), C.c_int]),
    "vt_decoder_grad_bind": (C.c_int, [_P, C.c_char_p, _P]),
    "vt_decoder_backward": (C.c_int, [_P, C.POINTER(DecoderBackwardArgs)]),
    "vt_decoder_tape_release": (C.c_int, [_P, C.c_int]),
    "vt_embed_loss": (C.c_int, [_P, C.POINTER(EmbedLossArgs)]),
    "vt_mse_loss": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P]),
    "vt_adaptive_loss_weights": (C.c_int, [_P, _P, _P, C.c_int, C.c_float, _P, _P, _P, _P]),
    "vt_op_conv2d_backward": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 7 + [_P, _P, _P, _P]),
    "vt_op_group_norm_backward": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int] * 4 + [C.c_float, C.c_int, C.c_int, _P, _P, _P, _P]),
    "vt_op_resnet_block_backward": (C.c_int, [_P, _P, C.POINTER(ResnetBlockPtrs), _P] + [C.c_int] * 6 +
                                    [C.POINTER(ResnetBlockGrads), _P]),
}

_lib = None


def lib_path() -> str:
    # VT_B200_LIB: development override (A/B builds of the same ABI)
    return os.environ.get("VT_B200_LIB") or _build.LIB_PATH


def load_library() -> C.CDLL:
    """Load the C-ABI library (no CUDA call is made).  Raises NativeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(
            f"{path} is missing: the sm_100a extension is not built (run `python -m vae_tagger_b200._build`); "
            "there is no CPU or PyTorch fallback for the encode+tag path")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.vt_abi_version() != 1:
        raise NativeError(f"ABI version mismatch: library {lib.vt_abi_version()}, binding 1")
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load_library().vt_last_error()
        raise NativeError(f"vae_tagger_b200 native call failed ({rc}): {msg.decode(errors='replace') if msg else '?'}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def smart_crop_box(src_w: int, src_h: int, dst_w: int, dst_h: int):
    """SmartResize's centre crop box (left, top, right, bottom) -- host-only, no GPU needed."""
    box = (C.c_int32 * 4)()
    _check(load_library().vt_smart_crop_box(src_w, src_h, dst_w, dst_h, box))
    return tuple(box)


def resize_coefficients(in_size: int, out_size: int, filter: int = FILTER_LANCZOS):
    """(ksize, bounds [out,2], kk [out,ksize]) int32 numpy arrays of one resize axis -- host-only."""
    import numpy as np

    lib = load_library()
    ks = C.c_int32()
    _check(lib.vt_resize_coefficients(in_size, out_size, filter, C.byref(ks), None, None))
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ks.value), np.int32)
    _check(lib.vt_resize_coefficients(in_size, out_size, filter, C.byref(ks),
                                      bounds.ctypes.data_as(C.POINTER(C.c_int32)),
                                      kk.ctypes.data_as(C.POINTER(C.c_int32))))
    return ks.value, bounds, kk


class Context:
    """One native context per CUDA device (per rank)."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise NativeError("vae_tagger_b200 needs a CUDA device (B200, sm_100a); torch.cuda.is_available() is False")
        self.lib = load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
            h = C.c_void_p()
            _check(self.lib.vt_ctx_create(self.device.index, C.byref(h)))
        self.h = h
        self.latent_channels = 16
        self.num_blocks = 4
        self.num_classes = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.vt_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ encoder
    def configure_encoder(self, cfg: dict):
        ec = EncoderConfig()
        chans = list(cfg.get("block_out_channels", [128, 256, 512, 512]))
        ec.in_channels = int(cfg.get("in_channels", 3))
        ec.num_blocks = len(chans)
        for i, ch in enumerate(chans):
            ec.block_out_channels[i] = int(ch)
        ec.layers_per_block = int(cfg.get("layers_per_block", 2))
        ec.norm_num_groups = int(cfg.get("norm_num_groups", 32))
        ec.latent_channels = int(cfg.get("latent_channels", 16))
        ec.mid_block_add_attention = int(bool(cfg.get("mid_block_add_attention", True)))
        ec.has_scaling_factor = int(cfg.get("scaling_factor") is not None)
        ec.has_shift_factor = int(cfg.get("shift_factor") is not None)
        ec.scaling_factor = float(cfg.get("scaling_factor") or 1.0)
        ec.shift_factor = float(cfg.get("shift_factor") or 0.0)
        _check(self.lib.vt_encoder_configure(self.h, C.byref(ec)))
        self.latent_channels = ec.latent_channels
        self.num_blocks = ec.num_blocks

    def _set_params(self, fn, sd: Dict[str, torch.Tensor]):
        for name, t in sd.items():
            if not torch.is_floating_point(t):
                continue
            t32 = t.detach().to(dtype=torch.float32).contiguous()
            shape = (C.c_int64 * max(1, t32.dim()))(*t32.shape)
            _check(fn(self.h, name.encode(), C.c_void_p(t32.data_ptr()), shape, t32.dim()))

    def load_encoder(self, state_dict: Dict[str, torch.Tensor]):
        """state_dict keys relative to ``encoder.`` (diffusers naming)."""
        self._set_params(self.lib.vt_encoder_set_param, state_dict)
        _check(self.lib.vt_encoder_finalize(self.h))

    def encode(self, images: torch.Tensor, precision=PREC_BF16, sample=False, apply_scale_shift=True, seed=0,
               noise: Optional[torch.Tensor] = None, want_moments=False, micro_batch=0, single_lane=False):
        """images: [B,3,H,W] float (any float dtype; normalised) or [B,H,W,3] uint8, on this device.
        Returns latent [B,LC,H/8,W/8] fp32 (and mean, logvar when want_moments)."""
        if images.dtype == torch.uint8:
            fmt = IN_U8_NHWC
            assert images.dim() == 4 and images.shape[-1] == 3, "uint8 images must be [B,H,W,3]"
            x = images.to(self.device).contiguous()
            B, H, W = x.shape[0], x.shape[1], x.shape[2]
        else:
            fmt = IN_F32_NCHW
            assert images.dim() == 4 and images.shape[1] == 3, "float images must be [B,3,H,W]"
            x = _f32c(images, self.device)
            B, H, W = x.shape[0], x.shape[2], x.shape[3]
        down = 1 << (self.num_blocks - 1)
        lc = self.latent_channels
        lat = torch.empty(B, lc, H // down, W // down, device=self.device, dtype=torch.float32)
        mean = torch.empty_like(lat) if want_moments else None
        logvar = torch.empty_like(lat) if want_moments else None
        nz = _f32c(noise, self.device) if noise is not None else None
        if B == 0:  # an empty batch is an empty result, as in PyTorch
            return (lat, mean, logvar) if want_moments else lat
        a = EncodeArgs()
        a.images = x.data_ptr(); a.in_fmt = fmt; a.batch = B; a.height = H; a.width = W
        a.precision = precision; a.sample = int(bool(sample)); a.apply_scale_shift = int(bool(apply_scale_shift))
        a.seed = int(seed) & (2**64 - 1)
        a.noise = nz.data_ptr() if nz is not None else None
        a.latent = lat.data_ptr()
        a.mean = mean.data_ptr() if mean is not None else None
        a.logvar = logvar.data_ptr() if logvar is not None else None
        a.micro_batch = int(micro_batch)
        a.single_lane = int(bool(single_lane))
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_encode(self.h, C.byref(a)))
        if want_moments:
            return lat, mean, logvar
        return lat

    # ------------------------------------------------------------------ encoder training
    def encode_train(self, images: torch.Tensor, precision=PREC_BF16, slot=0):
        """Training forward: ``(mean, logvar)`` of the posterior, with every activation kept on tape ``slot`` for
        ``encoder_backward``.  Image sizes must be multiples of 8."""
        if images.dtype == torch.uint8:
            fmt = IN_U8_NHWC
            x = images.to(self.device).contiguous()
            B, H, W = x.shape[0], x.shape[1], x.shape[2]
        else:
            fmt = IN_F32_NCHW
            x = _f32c(images, self.device)
            B, H, W = x.shape[0], x.shape[2], x.shape[3]
        down = 1 << (self.num_blocks - 1)
        mean = torch.empty(B, self.latent_channels, H // down, W // down, device=self.device, dtype=torch.float32)
        logvar = torch.empty_like(mean)
        a = EncodeArgs()
        a.images = x.data_ptr(); a.in_fmt = fmt; a.batch = B; a.height = H; a.width = W
        a.precision = precision; a.sample = 0; a.apply_scale_shift = 0
        a.mean = mean.data_ptr(); a.logvar = logvar.data_ptr(); a.latent = None
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_encoder_train_forward(self.h, C.byref(a), int(slot)))
        self._tape_inputs = getattr(self, "_tape_inputs", {})
        self._tape_inputs[slot] = x      # the backward reads the image again (conv_in weight gradient)
        return mean, logvar

    def encoder_backward(self, grad_mean, grad_logvar, grads: Dict[str, torch.Tensor], slot=0, accumulate=False):
        """Back-propagate through the forward kept on ``slot``.  ``grads``: fp32 device tensor per encoder parameter
        name (diffusers key without the ``encoder.`` prefix), written -- or added to when ``accumulate``."""
        gm = None if grad_mean is None else _f32c(grad_mean, self.device)
        gl = None if grad_logvar is None else _f32c(grad_logvar, self.device)
        for name, g in grads.items():
            assert g.dtype == torch.float32 and g.is_contiguous() and g.device.index == self.device.index
            _check(self.lib.vt_encoder_grad_bind(self.h, name.encode(), g.data_ptr()))
        a = EncoderBackwardArgs()
        a.grad_mean = gm.data_ptr() if gm is not None else None
        a.grad_logvar = gl.data_ptr() if gl is not None else None
        a.slot, a.accumulate = int(slot), int(bool(accumulate))
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_encoder_backward(self.h, C.byref(a)))

    def release_tape(self, slot=0):
        _check(self.lib.vt_encoder_tape_release(self.h, int(slot)))
        getattr(self, "_tape_inputs", {}).pop(slot, None)

    # ------------------------------------------------------------------ VAE decoder
    def load_decoder(self, state_dict: Dict[str, torch.Tensor]):
        """state_dict keys relative to ``decoder.`` (diffusers naming); needs ``configure_encoder`` first."""
        self._set_params(self.lib.vt_decoder_set_param, state_dict)
        _check(self.lib.vt_decoder_finalize(self.h))

    def decode(self, latent: torch.Tensor, precision=PREC_BF16, apply_scale_shift=False, micro_batch=0):
        """latent [B,LC,h,w] -> image [B,3,8h,8w] fp32 (``vae.decode(z).sample``; with apply_scale_shift the
        ``(z - shift) / scale`` of ``DiffusersVAEWrapper.decode`` first)."""
        lat = _f32c(latent, self.device)
        B, _, h, w = lat.shape
        up = 1 << (self.num_blocks - 1)
        img = torch.empty(B, 3, h * up, w * up, device=self.device, dtype=torch.float32)
        if B == 0:
            return img
        a = DecodeArgs()
        a.latent = lat.data_ptr(); a.batch = B; a.lat_h = h; a.lat_w = w; a.precision = precision
        a.apply_scale_shift = int(bool(apply_scale_shift)); a.image = img.data_ptr()
        a.micro_batch = int(micro_batch)
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_decode(self.h, C.byref(a)))
        return img

    def decode_train(self, latent: torch.Tensor, precision=PREC_BF16, apply_scale_shift=False, slot=0):
        """Training forward of the decoder: the image, with every activation kept on decoder tape ``slot``."""
        lat = _f32c(latent, self.device)
        B, _, h, w = lat.shape
        up = 1 << (self.num_blocks - 1)
        img = torch.empty(B, 3, h * up, w * up, device=self.device, dtype=torch.float32)
        a = DecodeArgs()
        a.latent = lat.data_ptr(); a.batch = B; a.lat_h = h; a.lat_w = w; a.precision = precision
        a.apply_scale_shift = int(bool(apply_scale_shift)); a.image = img.data_ptr()
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_decoder_train_forward(self.h, C.byref(a), int(slot)))
        return img

    def decoder_backward(self, grad_image, grads: Dict[str, torch.Tensor], want_latent_grad=True, slot=0, accumulate=False):
        """Back-propagate ``d loss / d image`` through the decoder forward kept on ``slot``: fills ``grads`` (fp32 tensor
        per decoder parameter name, diffusers key without ``decoder.``) and returns ``d loss / d latent`` (or None)."""
        gi = _f32c(grad_image, self.device)
        for name, g in grads.items():
            assert g.dtype == torch.float32 and g.is_contiguous() and g.device.index == self.device.index
            _check(self.lib.vt_decoder_grad_bind(self.h, name.encode(), g.data_ptr()))
        B, _, H, W = gi.shape
        down = 1 << (self.num_blocks - 1)
        gz = torch.empty(B, self.latent_channels, H // down, W // down, device=self.device) if want_latent_grad else None
        a = DecoderBackwardArgs()
        a.grad_image = gi.data_ptr(); a.grad_latent = gz.data_ptr() if gz is not None else None
        a.slot, a.accumulate = int(slot), int(bool(accumulate))
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_decoder_backward(self.h, C.byref(a)))
        return gz

    def release_decoder_tape(self, slot=0):
        _check(self.lib.vt_decoder_tape_release(self.h, int(slot)))

    # ------------------------------------------------------------------ head
    def configure_head(self, kind: int, latent_channels: int, num_classes: int, use_spatial_attention=True,
                       use_self_attention=True, attention_heads=8, use_cross_attention=False, plain_flat_dim=0):
        hc = HeadConfig(kind, latent_channels, num_classes, int(use_spatial_attention), int(use_self_attention),
                        attention_heads, int(use_cross_attention), int(plain_flat_dim))
        _check(self.lib.vt_head_configure(self.h, C.byref(hc)))
        self.num_classes = num_classes

    def load_head(self, state_dict: Dict[str, torch.Tensor]):
        self._set_params(self.lib.vt_head_set_param, state_dict)
        _check(self.lib.vt_head_finalize(self.h))

    def tag(self, latent: torch.Tensor, threshold=0.5, want=("logits", "probs", "conf", "idx", "count")):
        lat = _f32c(latent, self.device)
        B, _, h, w = lat.shape
        T = self.num_classes
        out = {}
        if "logits" in want:
            out["logits"] = torch.empty(B, T, device=self.device, dtype=torch.float32)
        if "probs" in want:
            out["probs"] = torch.empty(B, T, device=self.device, dtype=torch.float32)
        if "conf" in want:
            out["conf"] = torch.empty(B, T, device=self.device, dtype=torch.float32)
        if "idx" in want:
            out["idx"] = torch.empty(B, T, device=self.device, dtype=torch.int64)
        if "count" in want:
            out["count"] = torch.empty(B, device=self.device, dtype=torch.int32)
        if B == 0:
            return out
        a = TagArgs()
        a.latent = lat.data_ptr(); a.batch = B; a.lat_h = h; a.lat_w = w; a.threshold = float(threshold)
        a.logits = out["logits"].data_ptr() if "logits" in out else None
        a.probs = out["probs"].data_ptr() if "probs" in out else None
        a.conf_sorted = out["conf"].data_ptr() if "conf" in out else None
        a.idx_sorted = out["idx"].data_ptr() if "idx" in out else None
        a.count = out["count"].data_ptr() if "count" in out else None
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_tag(self.h, C.byref(a)))
        return out

    def infer(self, images: torch.Tensor, threshold=0.5, precision=PREC_BF16, micro_batch=0, single_lane=False):
        """encode (mode, scale/shift) + tag on DEVICE tensors in one call: the head of each internal micro-batch
        runs right behind its encoder.  images: float [B,3,H,W] or uint8 [B,H,W,3] on this device.  Returns a
        dict of device tensors: conf, idx (sorted descending), count (conf >= threshold), latent."""
        if images.dtype == torch.uint8:
            fmt = IN_U8_NHWC
            x = images.to(self.device).contiguous()
            B, H, W = x.shape[0], x.shape[1], x.shape[2]
        else:
            fmt = IN_F32_NCHW
            x = _f32c(images, self.device)
            B, H, W = x.shape[0], x.shape[2], x.shape[3]
        T = self.num_classes
        down = 1 << (self.num_blocks - 1)
        out = {
            "conf": torch.empty(B, T, device=self.device, dtype=torch.float32),
            "idx": torch.empty(B, T, device=self.device, dtype=torch.int64),
            "count": torch.empty(B, device=self.device, dtype=torch.int32),
            "latent": torch.empty(B, self.latent_channels, H // down, W // down, device=self.device),
        }
        if B == 0:
            return out
        a = InferArgs()
        a.images = x.data_ptr(); a.in_fmt = fmt; a.batch = B; a.height = H; a.width = W
        a.precision = precision; a.threshold = float(threshold)
        a.conf_sorted = out["conf"].data_ptr(); a.idx_sorted = out["idx"].data_ptr()
        a.count = out["count"].data_ptr(); a.latent = out["latent"].data_ptr()
        a.micro_batch = int(micro_batch); a.single_lane = int(bool(single_lane))
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_infer(self.h, C.byref(a)))
        return out

    def infer_host(self, images_host: torch.Tensor, threshold=0.5, precision=PREC_BF16, want_latent=False,
                   micro_batch=0, out=None):
        """End-to-end on HOST tensors (pinned recommended): H2D + encode + tag + D2H, synchronous."""
        x = images_host
        assert x.device.type == "cpu" and x.is_contiguous()
        if x.dtype == torch.uint8:
            fmt, B, H, W = IN_U8_NHWC, x.shape[0], x.shape[1], x.shape[2]
        else:
            assert x.dtype == torch.float32
            fmt, B, H, W = IN_F32_NCHW, x.shape[0], x.shape[2], x.shape[3]
        T = self.num_classes
        down = 1 << (self.num_blocks - 1)
        if out is None:
            out = {
                "conf": torch.empty(B, T, dtype=torch.float32).pin_memory(),
                "idx": torch.empty(B, T, dtype=torch.int64).pin_memory(),
                "count": torch.empty(B, dtype=torch.int32).pin_memory(),
            }
            if want_latent:
                out["latent"] = torch.empty(B, self.latent_channels, H // down, W // down).pin_memory()
        a = InferHostArgs()
        a.images_host = x.data_ptr(); a.in_fmt = fmt; a.batch = B; a.height = H; a.width = W
        a.precision = precision; a.threshold = float(threshold)
        a.conf_sorted_host = out["conf"].data_ptr(); a.idx_sorted_host = out["idx"].data_ptr()
        a.count_host = out["count"].data_ptr()
        a.latent_host = out["latent"].data_ptr() if "latent" in out else None
        a.micro_batch = int(micro_batch)
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_infer_host(self.h, C.byref(a)))
        return out

    # ------------------------------------------------------------------ loss
    def focal_loss(self, logits: torch.Tensor, targets: torch.Tensor, alpha=1.0, gamma=2.0, want_grad=True):
        """Returns (sum of focal terms as a 1-element device tensor, d(mean)/d(logits) or None)."""
        x = _f32c(logits, self.device)
        y = _f32c(targets, self.device)
        n = x.numel()
        loss = torch.zeros(1, device=self.device, dtype=torch.float32)
        grad = torch.empty_like(x) if want_grad else None
        with torch.cuda.device(self.device):
            _check(self.lib.vt_focal_loss(self.h, _ptr(x), _ptr(y), n, float(alpha), float(gamma), 1.0 / n, _ptr(loss),
                                          _ptr(grad), _stream(self.device)))
        return loss, grad

    # ------------------------------------------------------------------ preprocessing
    def resize_u8(self, src: torch.Tensor, size, box=None, filter=FILTER_LANCZOS, out: torch.Tensor = None):
        """``PIL.Image.fromarray(src).crop(box).resize(size, filter)`` on the device, bit-exact.
        src: uint8 [h,w,3] CUDA tensor (row stride may exceed w*3); size = (W, H); box = (l,t,r,b) or None;
        out: optional uint8 [H,W,3] view to write into (e.g. one image of a batch buffer)."""
        assert src.dtype == torch.uint8 and src.is_cuda and src.dim() == 3 and src.shape[2] == 3
        assert src.stride(2) == 1 and src.stride(1) == 3, "pixels must be packed RGB"
        W, H = int(size[0]), int(size[1])
        if out is None:
            out = torch.empty(H, W, 3, dtype=torch.uint8, device=self.device)
        assert out.dtype == torch.uint8 and tuple(out.shape) == (H, W, 3) and out.stride(2) == 1 and out.stride(1) == 3
        h, w = src.shape[0], src.shape[1]
        l, t, r, b = box if box is not None else (0, 0, w, h)
        a = ResizeArgs()
        a.src = src.data_ptr(); a.src_w = w; a.src_h = h; a.src_stride = src.stride(0)
        a.crop_l, a.crop_t, a.crop_r, a.crop_b = int(l), int(t), int(r), int(b)
        a.dst = out.data_ptr(); a.dst_w = W; a.dst_h = H; a.dst_stride = out.stride(0)
        a.filter = int(filter)
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_resize_u8(self.h, C.byref(a)))
        return out

    def resize_u8_batch(self, srcs, size, boxes=None, filter=FILTER_LANCZOS, out: torch.Tensor = None):
        """``resize_u8`` for a list of source images into one ``[N, H, W, 3]`` uint8 batch (one native call)."""
        W, H = int(size[0]), int(size[1])
        n = len(srcs)
        if out is None:
            out = torch.empty(n, H, W, 3, dtype=torch.uint8, device=self.device)
        assert out.dtype == torch.uint8 and tuple(out.shape) == (n, H, W, 3) and out.is_contiguous()
        items = (ResizeArgs * n)()
        for i, src in enumerate(srcs):
            assert src.dtype == torch.uint8 and src.is_cuda and src.dim() == 3 and src.shape[2] == 3
            assert src.stride(2) == 1 and src.stride(1) == 3, "pixels must be packed RGB"
            h, w = src.shape[0], src.shape[1]
            l, t, r, b = boxes[i] if boxes is not None and boxes[i] is not None else (0, 0, w, h)
            a = items[i]
            a.src = src.data_ptr(); a.src_w = w; a.src_h = h; a.src_stride = src.stride(0)
            a.crop_l, a.crop_t, a.crop_r, a.crop_b = int(l), int(t), int(r), int(b)
            a.dst = out[i].data_ptr(); a.dst_w = W; a.dst_h = H; a.dst_stride = out.stride(1)
            a.filter = int(filter)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_resize_u8_batch(self.h, items, n, _stream(self.device)))
        return out

    # ------------------------------------------------------------------ head training step
    def head_param_layout(self):
        """[(state-dict key, offset, numel)] of the flat parameter / gradient buffers, in the order of the
        reference module's ``parameters()``; needs ``configure_head`` first."""
        n, total = C.c_int32(), C.c_int64()
        _check(self.lib.vt_head_param_count(self.h, C.byref(n), C.byref(total)))
        out = []
        buf = C.create_string_buffer(128)
        for i in range(n.value):
            off, numel = C.c_int64(), C.c_int64()
            _check(self.lib.vt_head_param_layout(self.h, i, buf, 128, C.byref(off), C.byref(numel)))
            out.append((buf.value.decode(), off.value, numel.value))
        assert not out or out[-1][1] + out[-1][2] == total.value
        return out

    def head_train_step(self, latent, targets, flat_params, flat_grads=None, bn_running_mean=None,
                        bn_running_var=None, bn_num_batches_tracked=None, bn_momentum=0.1, focal_alpha=1.0,
                        focal_gamma=2.0, loss_scale=1.0, dropout=True, attention_dropout=0.1, seed=0, loss=None,
                        want_logits=False, class_weights=None):
        """Train-mode forward + focal/BCE loss + backward of the configured head (``vt_head_train_step``).
        Gradients are accumulated into ``flat_grads``; returns (loss accumulator tensor, logits or None)."""
        lat = _f32c(latent, self.device)
        tgt = _f32c(targets, self.device)
        B, _, h, w = lat.shape
        for t in (flat_params, flat_grads, bn_running_mean, bn_running_var):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
        assert bn_num_batches_tracked is None or bn_num_batches_tracked.dtype == torch.int64
        if loss is None:
            loss = torch.zeros(1, device=self.device, dtype=torch.float32)
        logits = torch.empty(B, self.num_classes, device=self.device, dtype=torch.float32) if want_logits else None
        a = HeadTrainArgs()
        a.latent = lat.data_ptr(); a.targets = tgt.data_ptr(); a.batch = B; a.lat_h = h; a.lat_w = w
        a.params = flat_params.data_ptr()
        a.grads = flat_grads.data_ptr() if flat_grads is not None else None
        a.bn_running_mean = bn_running_mean.data_ptr() if bn_running_mean is not None else None
        a.bn_running_var = bn_running_var.data_ptr() if bn_running_var is not None else None
        a.bn_num_batches_tracked = bn_num_batches_tracked.data_ptr() if bn_num_batches_tracked is not None else None
        a.bn_momentum = float(bn_momentum); a.focal_alpha = float(focal_alpha); a.focal_gamma = float(focal_gamma)
        a.loss_scale = float(loss_scale); a.dropout = int(bool(dropout))
        a.attention_dropout = float(attention_dropout); a.seed = int(seed) & (2 ** 64 - 1)
        a.loss = loss.data_ptr(); a.logits = logits.data_ptr() if logits is not None else None
        if class_weights is not None:
            assert class_weights.is_cuda and class_weights.dtype == torch.float32 and class_weights.is_contiguous()
            assert class_weights.numel() == self.num_classes
            a.class_weights = class_weights.data_ptr()
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_head_train_step(self.h, C.byref(a)))
        return loss, logits

    def head_dropout_masks(self, batch, attention_dropout, seed, heads=8, widths=(1024, 512, 256)):
        """The keep-masks (0/1) ``head_train_step`` uses for ``seed``: (attn [B,heads,64,64], [cls_i [B,width_i]])."""
        attn = torch.empty(batch, heads, 64, 64, device=self.device, dtype=torch.float32)
        cls = [torch.empty(batch, w, device=self.device, dtype=torch.float32) for w in widths]
        ptrs = [_ptr(c) for c in cls] + [None] * (3 - len(cls))
        with torch.cuda.device(self.device):
            _check(self.lib.vt_head_dropout_masks(self.h, batch, float(attention_dropout), int(seed) & (2 ** 64 - 1),
                                                  _ptr(attn), ptrs[0], ptrs[1], ptrs[2], _stream(self.device)))
        return attn, cls

    def adamw_step(self, params, grads, exp_avg, exp_avg_sq, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                   step=1, grad_scale=1.0, max_norm=0.0, zero_grad=True, norm_out=None):
        """clip_grad_norm_ + AdamW + zero_grad on flat fp32 device buffers (``vt_adamw_step``)."""
        for t in (params, grads, exp_avg, exp_avg_sq):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()
        with torch.cuda.device(self.device):
            _check(self.lib.vt_adamw_step(self.h, _ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq),
                                          params.numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                          float(weight_decay), int(step), float(grad_scale), float(max_norm),
                                          int(bool(zero_grad)), _ptr(norm_out), _stream(self.device)))

    # ------------------------------------------------------------------ accounting
    def profile_enable(self, timing: bool):
        _check(self.lib.vt_profile_enable(self.h, int(timing)))

    def profile_read(self, reset=True):
        buf = (C.c_double * (NUM_KERNEL_CLASSES * 4))()
        _check(self.lib.vt_profile_read(self.h, buf, int(reset)))
        return {KERNEL_CLASS_NAMES[i]: {"launches": buf[4 * i], "ms": buf[4 * i + 1], "flops": buf[4 * i + 2],
                                        "bytes": buf[4 * i + 3]} for i in range(NUM_KERNEL_CLASSES)}

    # ------------------------------------------------------------------ single ops (tests)
    def op_conv2d(self, x, w, bias=None, residual=None, sc_x=None, sc_w=None, stride=1, precision=PREC_BF16,
                  want_stats=False):
        x = _f32c(x, self.device); w = _f32c(w, self.device)
        bias = _f32c(bias, self.device) if bias is not None else None
        residual = _f32c(residual, self.device) if residual is not None else None
        sc_x = _f32c(sc_x, self.device) if sc_x is not None else None
        sc_w = _f32c(sc_w, self.device) if sc_w is not None else None
        N, Cin, H, W = x.shape
        Cout, _, k, _ = w.shape
        Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
        out = torch.empty(N, Cout, Ho, Wo, device=self.device, dtype=torch.float32)
        stats = torch.zeros(N, 32, 2, device=self.device, dtype=torch.float64) if want_stats else None
        Cs = sc_x.shape[1] if sc_x is not None else 0
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_conv2d(self.h, _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), _ptr(sc_x), _ptr(sc_w),
                                         N, Cin, H, W, Cout, k, stride, Cs, precision, _ptr(out), _ptr(stats),
                                         _stream(self.device)))
        return (out, stats) if want_stats else out

    def op_conv3_fused(self, x, gamma, beta, w, bias=None, residual=None, eps=1e-6, silu=True, want_stats=False,
                       sc_x=None, sc_w=None):
        x = _f32c(x, self.device); w = _f32c(w, self.device)
        gamma = _f32c(gamma, self.device); beta = _f32c(beta, self.device)
        bias = _f32c(bias, self.device) if bias is not None else None
        residual = _f32c(residual, self.device) if residual is not None else None
        sc_x = _f32c(sc_x, self.device) if sc_x is not None else None
        sc_w = _f32c(sc_w, self.device) if sc_w is not None else None
        N, Cin, H, W = x.shape
        Cout = w.shape[0]
        Cs = sc_x.shape[1] if sc_x is not None else 0
        out = torch.empty(N, Cout, H, W, device=self.device, dtype=torch.float32)
        stats = torch.zeros(N, 32, 2, device=self.device, dtype=torch.float64) if want_stats else None
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_conv3_fused(self.h, _ptr(x), _ptr(gamma), _ptr(beta), _ptr(w), _ptr(bias),
                                              _ptr(residual), _ptr(sc_x), _ptr(sc_w), N, Cin, H, W, Cout, Cs,
                                              float(eps), int(silu), _ptr(out), _ptr(stats), _stream(self.device)))
        return (out, stats) if want_stats else out

    def op_flash_attention(self, q, k, v, bias_v=None, scale=None):
        """q, k, v: [n, tokens, 512] fp32 -> softmax(scale q k^T) v + bias_v, [n, tokens, 512] fp32."""
        q = _f32c(q, self.device); k = _f32c(k, self.device); v = _f32c(v, self.device)
        n, tokens, d = q.shape
        qk = torch.cat([q, k], dim=2).contiguous()
        vt = v.transpose(1, 2).contiguous()
        bias_v = _f32c(bias_v, self.device) if bias_v is not None else None
        out = torch.empty(n, tokens, d, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_flash_attention(self.h, _ptr(qk), _ptr(vt), _ptr(bias_v), n, tokens,
                                                  float(scale if scale is not None else d ** -0.5), _ptr(out),
                                                  _stream(self.device)))
        return out

    def op_gemm_nt(self, A, B, bias=None, alpha=1.0, precision=PREC_BF16):
        A = _f32c(A, self.device); B = _f32c(B, self.device)
        bias = _f32c(bias, self.device) if bias is not None else None
        if A.dim() == 2:
            A = A[None]
        batch, M, K = A.shape
        b_batched = int(B.dim() == 3)
        N = B.shape[-2]
        out = torch.empty(batch, M, N, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_gemm_nt(self.h, _ptr(A), _ptr(B), _ptr(bias), batch, M, N, K, b_batched, float(alpha),
                                          precision, _ptr(out), _stream(self.device)))
        return out

    def op_group_norm(self, x, gamma, beta, groups=32, eps=1e-6, silu=False, precision=PREC_BF16):
        x = _f32c(x, self.device); gamma = _f32c(gamma, self.device); beta = _f32c(beta, self.device)
        N, Cc, H, W = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_group_norm(self.h, _ptr(x), _ptr(gamma), _ptr(beta), N, Cc, H, W, groups, float(eps),
                                             int(silu), precision, _ptr(out), _stream(self.device)))
        return out

    def op_softmax_rows(self, s, precision=PREC_BF16):
        s = _f32c(s, self.device)
        rows, cols = s.shape
        out = torch.empty_like(s)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_softmax_rows(self.h, _ptr(s), rows, cols, precision, _ptr(out), _stream(self.device)))
        return out


    # ---- VAE fine-tuning losses (improved_losses.py): value + analytic gradients in one call
    def embed_loss(self, kind, a, p, n=None, labels_a=None, labels_p=None, margin=1.0, similarity="cosine",
                   want_grad=True):
        """``kind`` 0: ImprovedTripletLoss(a, p, n, labels_a, labels_p); 1: ContrastiveLoss(a, p, labels_a, labels_p).
        Returns ``(loss [1], (grad_a, grad_p, grad_n))`` with ``d loss / d input`` (``None`` when not wanted)."""
        a = _f32c(a, self.device); p = _f32c(p, self.device)
        n = None if n is None else _f32c(n, self.device)
        la = None if labels_a is None else _f32c(labels_a, self.device)
        lp = None if labels_p is None else _f32c(labels_p, self.device)
        B, D = a.shape
        args = EmbedLossArgs()
        loss = torch.empty(1, device=self.device)
        ga = torch.empty_like(a) if want_grad else None
        gp = torch.empty_like(p) if want_grad else None
        gn = torch.empty_like(n) if (want_grad and n is not None and kind == 0) else None
        args.a, args.p, args.n = a.data_ptr(), p.data_ptr(), (n.data_ptr() if n is not None else None)
        args.labels_a = la.data_ptr() if la is not None else None
        args.labels_p = lp.data_ptr() if lp is not None else None
        args.B, args.D, args.T = B, D, (la.shape[1] if la is not None else 0)
        args.kind, args.similarity, args.margin = kind, (0 if similarity == "cosine" else 1), float(margin)
        args.loss = loss.data_ptr()
        args.grad_a = ga.data_ptr() if ga is not None else None
        args.grad_p = gp.data_ptr() if gp is not None else None
        args.grad_n = gn.data_ptr() if gn is not None else None
        args.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.vt_embed_loss(self.h, C.byref(args)))
        return loss, (ga, gp, gn)

    def mse_loss(self, x, y, want_grad=True):
        x = _f32c(x, self.device); y = _f32c(y, self.device)
        loss = torch.empty(1, device=self.device)
        gx = torch.empty_like(x) if want_grad else None
        with torch.cuda.device(self.device):
            _check(self.lib.vt_mse_loss(self.h, _ptr(x), _ptr(y), x.numel(), _ptr(loss), _ptr(gx), _stream(self.device)))
        return loss, gx

    def adaptive_loss_weights(self, log_w, losses, temperature=1.0):
        log_w = _f32c(log_w, self.device); losses = _f32c(losses, self.device)
        total = torch.empty(1, device=self.device)
        weights, glw = torch.empty_like(log_w), torch.empty_like(log_w)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_adaptive_loss_weights(self.h, _ptr(log_w), _ptr(losses), log_w.numel(), float(temperature),
                                                     _ptr(total), _ptr(weights), _ptr(glw), _stream(self.device)))
        return total, weights, glw

    # ---- backward building blocks of the encoder
    def op_conv2d_backward(self, x, w, grad_out, precision=PREC_F16, want=(True, True, True)):
        """(grad_x, grad_w, grad_b) of ``conv2d(x, w, b, padding=k//2)`` (stride 1, k = 1 or 3)."""
        x = _f32c(x, self.device); w = _f32c(w, self.device); go = _f32c(grad_out, self.device)
        N, Cin, H, W = x.shape
        Cout, k = w.shape[0], w.shape[-1]
        gx = torch.empty_like(x) if want[0] else None
        gw = torch.empty_like(w) if want[1] else None
        gb = torch.empty(Cout, device=self.device) if want[2] else None
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_conv2d_backward(self.h, _ptr(x), _ptr(w), _ptr(go), N, Cin, H, W, Cout, k, precision,
                                                  _ptr(gx), _ptr(gw), _ptr(gb), _stream(self.device)))
        return gx, gw, gb

    def op_group_norm_backward(self, x, gamma, beta, grad_y, eps=1e-6, silu=True, precision=PREC_F16):
        """(grad_x, grad_gamma, grad_beta) of ``act(group_norm(x, 32, gamma, beta))``."""
        x = _f32c(x, self.device); gamma = _f32c(gamma, self.device); beta = _f32c(beta, self.device)
        gy = _f32c(grad_y, self.device)
        N, Cc, H, W = x.shape
        gx, gg, gb = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(beta)
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_group_norm_backward(self.h, _ptr(x), _ptr(gamma), _ptr(beta), _ptr(gy), N, Cc, H, W,
                                                      float(eps), int(silu), precision, _ptr(gx), _ptr(gg), _ptr(gb),
                                                      _stream(self.device)))
        return gx, gg, gb

    def op_resnet_block_backward(self, x, params, grad_out, precision=PREC_F16):
        """Backward of diffusers' ResnetBlock2D.  ``params``: dict with the block's state-dict keys (``norm1.weight``,
        ``conv1.weight``, ..., optional ``conv_shortcut.weight/bias``).  Returns ``(grad_x, {key: grad})``."""
        x = _f32c(x, self.device); go = _f32c(grad_out, self.device)
        keys = {"norm1_w": "norm1.weight", "norm1_b": "norm1.bias", "conv1_w": "conv1.weight", "conv1_b": "conv1.bias",
                "norm2_w": "norm2.weight", "norm2_b": "norm2.bias", "conv2_w": "conv2.weight", "conv2_b": "conv2.bias",
                "sc_w": "conv_shortcut.weight", "sc_b": "conv_shortcut.bias"}
        N, Cin, H, W = x.shape
        Cout = params["conv1.weight"].shape[0]
        pp, gp, held, grads = ResnetBlockPtrs(), ResnetBlockGrads(), [], {}
        for f, k in keys.items():
            if k in params:
                t = _f32c(params[k], self.device)
                g = torch.empty_like(t)
                held.append(t)
                grads[k] = g
                setattr(pp, f, t.data_ptr()); setattr(gp, f, g.data_ptr())
        gx = torch.empty_like(x)
        gp.x = gx.data_ptr()
        with torch.cuda.device(self.device):
            _check(self.lib.vt_op_resnet_block_backward(self.h, _ptr(x), C.byref(pp), _ptr(go), N, Cin, Cout, H, W,
                                                        precision, C.byref(gp), _stream(self.device)))
        return gx, grads


_contexts: Dict[int, Context] = {}


def get_context(device=None) -> Context:
    """Process-wide context of a device (created on first use)."""
    if not torch.cuda.is_available():
        raise NativeError("vae_tagger_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    if idx not in _contexts:
        _contexts[idx] = Context(torch.device("cuda", idx))
    return _contexts[idx]
