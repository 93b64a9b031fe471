"""``AutoencoderKL`` -- the object the reference builds at ``diffusers_vae_loader.py:8-35`` --
backed by the sm_100a encoder kernels.

Only what the encode+tag hot path touches is implemented:
``.config`` (attribute access, incl. ``scaling_factor`` / ``shift_factor``),
``.encode(x).latent_dist`` with ``.mode() / .sample() / .kl() / .mean / .logvar / .std / .var``,
``.load_state_dict(sd, strict=False)`` (diffusers key names, Appendix B), ``.parameters()``, ``.to()``,
``.eval()``.  ``.decode(z).sample`` / ``.forward`` (the VAE decoder) run the same kernel
family; the decoder's parameters are materialised on demand (a checkpoint with ``decoder.*`` keys,
``decode()``, or ``enable_decoder()``) so that encode-only users do not pay for them.

The module holds ordinary ``nn.Parameter``s under diffusers' names so checkpoints load with the
stock ``nn.Module`` machinery; the parameters are mirrored into the native context (repacked to
``[Cout][kh][kw][Cin]`` bf16) lazily and again whenever they change.  There is no PyTorch
forward: every ``encode`` runs the CUDA path or raises.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import _native

_LEGACY_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


class _Config(SimpleNamespace):
    """Attribute + mapping access like diffusers' FrozenDict config."""

    def get(self, k, default=None):
        return getattr(self, k, default)

    def __getitem__(self, k):
        return getattr(self, k)

    def keys(self):
        return vars(self).keys()


def _resnet(cin, cout, groups):
    m = nn.Module()
    m.norm1 = nn.GroupNorm(groups, cin, eps=1e-6, affine=True)
    m.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
    m.norm2 = nn.GroupNorm(groups, cout, eps=1e-6, affine=True)
    m.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
    if cin != cout:
        m.conv_shortcut = nn.Conv2d(cin, cout, 1, 1, 0)
    return m


def _encoder_params(cfg) -> nn.Module:
    """Parameter containers with diffusers' ``Encoder`` state-dict layout (never called)."""
    chans = list(cfg.block_out_channels)
    g = cfg.norm_num_groups
    enc = nn.Module()
    enc.conv_in = nn.Conv2d(cfg.in_channels, chans[0], 3, 1, 1)
    blocks, cin = [], chans[0]
    for i, cout in enumerate(chans):
        b = nn.Module()
        b.resnets = nn.ModuleList([_resnet(cin if j == 0 else cout, cout, g) for j in range(cfg.layers_per_block)])
        if i < len(chans) - 1:
            d = nn.Module()
            d.conv = nn.Conv2d(cout, cout, 3, 2, 0)
            b.downsamplers = nn.ModuleList([d])
        blocks.append(b)
        cin = cout
    enc.down_blocks = nn.ModuleList(blocks)
    mid = nn.Module()
    c = chans[-1]
    if cfg.mid_block_add_attention:
        a = nn.Module()
        a.group_norm = nn.GroupNorm(g, c, eps=1e-6, affine=True)
        a.to_q = nn.Linear(c, c)
        a.to_k = nn.Linear(c, c)
        a.to_v = nn.Linear(c, c)
        a.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])
        mid.attentions = nn.ModuleList([a])
    mid.resnets = nn.ModuleList([_resnet(c, c, g), _resnet(c, c, g)])
    enc.mid_block = mid
    enc.conv_norm_out = nn.GroupNorm(g, c, eps=1e-6, affine=True)
    enc.conv_out = nn.Conv2d(c, 2 * cfg.latent_channels, 3, 1, 1)
    return enc


def _decoder_params(cfg) -> nn.Module:
    """Parameter containers with diffusers' ``Decoder`` state-dict layout (never called)."""
    chans = list(reversed(cfg.block_out_channels))
    g = cfg.norm_num_groups
    dec = nn.Module()
    dec.conv_in = nn.Conv2d(cfg.latent_channels, chans[0], 3, 1, 1)
    mid = nn.Module()
    c = chans[0]
    if cfg.mid_block_add_attention:
        a = nn.Module()
        a.group_norm = nn.GroupNorm(g, c, eps=1e-6, affine=True)
        a.to_q = nn.Linear(c, c)
        a.to_k = nn.Linear(c, c)
        a.to_v = nn.Linear(c, c)
        a.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])
        mid.attentions = nn.ModuleList([a])
    mid.resnets = nn.ModuleList([_resnet(c, c, g), _resnet(c, c, g)])
    dec.mid_block = mid
    blocks, cin = [], chans[0]
    for i, cout in enumerate(chans):
        b = nn.Module()
        b.resnets = nn.ModuleList([_resnet(cin if j == 0 else cout, cout, g) for j in range(cfg.layers_per_block + 1)])
        if i < len(chans) - 1:
            u = nn.Module()
            u.conv = nn.Conv2d(cout, cout, 3, 1, 1)
            b.upsamplers = nn.ModuleList([u])
        blocks.append(b)
        cin = cout
    dec.up_blocks = nn.ModuleList(blocks)
    dec.conv_norm_out = nn.GroupNorm(g, chans[-1], eps=1e-6, affine=True)
    dec.conv_out = nn.Conv2d(chans[-1], cfg.out_channels, 3, 1, 1)
    return dec


class DecoderOutput(SimpleNamespace):
    pass


class DiagonalGaussianDistribution:
    """Posterior returned by ``AutoencoderKL.encode(x).latent_dist``."""

    def __init__(self, mean: torch.Tensor, logvar: torch.Tensor):
        self.mean = mean
        self.logvar = logvar  # already clamped to [-30, 20] by the kernel
        self.deterministic = False

    @property
    def std(self):
        return torch.exp(0.5 * self.logvar)

    @property
    def var(self):
        return torch.exp(self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def kl(self, other=None):
        if other is not None:
            raise NotImplementedError("kl() against another distribution is not on the encode+tag path")
        return 0.5 * torch.sum(self.mean.pow(2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])


class _TapeSlot:
    """Owns one tape slot of the native context for the lifetime of an autograd graph node: the slot goes back to the
    pool when the backward has consumed it OR when the graph is dropped without a backward (a forward in ``train()``
    mode whose loss is never back-propagated must not leak one of the context's ``MAX_TAPES`` slots)."""

    def __init__(self, vae, pool):
        self.vae, self.pool = vae, pool
        self.slot = vae._take_tape_slot(pool)

    def release(self):
        if self.slot is not None:
            self.vae._free_tape_slot(self.slot, self.pool)
            self.slot = None

    def __del__(self):
        self.release()


class _EncodeTrainFn(torch.autograd.Function):
    """``encoder(x) -> (mean, logvar)`` with the native training forward / backward (what autograd does
    for the reference in train_full.py:201-256 / train_vae.py:124-186).  The activations stay inside the native context
    on a tape slot until the backward has run; the backward returns one gradient per encoder parameter."""

    @staticmethod
    def forward(ctx, vae, x, *params):
        nctx = vae._sync_native(x.device)
        ctx.tape = _TapeSlot(vae, "_tape_slots")
        slot = ctx.tape.slot
        mean, logvar = nctx.encode_train(x, precision=vae._precision(), slot=slot)
        ctx.vae, ctx.nctx, ctx.slot = vae, nctx, slot
        ctx.names = [n for n, _ in vae.encoder.named_parameters()]
        ctx.save_for_backward(logvar)
        ctx.mark_non_differentiable()
        return mean, logvar

    @staticmethod
    def backward(ctx, g_mean, g_logvar):
        (logvar,) = ctx.saved_tensors
        if g_logvar is not None:
            # the kernel returns logvar clamped to [-30, 20] (diffusers DiagonalGaussianDistribution): no gradient
            # flows through a saturated entry
            g_logvar = g_logvar * ((logvar > -30.0) & (logvar < 20.0))
        params = dict(ctx.vae.encoder.named_parameters())
        grads = {n: torch.empty_like(params[n], dtype=torch.float32) for n in ctx.names}
        if g_mean is None and g_logvar is None:
            for g in grads.values():
                g.zero_()
        else:
            ctx.nctx.encoder_backward(g_mean, g_logvar, grads, slot=ctx.slot, accumulate=False)
        ctx.tape.release()
        return (None, None) + tuple(grads[n].to(params[n].dtype) for n in ctx.names)


class _DecodeTrainFn(torch.autograd.Function):
    """``decoder(z) -> image`` with the native training forward / backward (the reference back-propagates its
    reconstruction MSE through ``vae.decode``: train_vae.py:124-186, improved_losses.py:278).  The backward returns
    ``d loss / d z`` and one gradient per decoder parameter."""

    @staticmethod
    def forward(ctx, vae, apply_scale_shift, z, *params):
        nctx = vae._sync_native_decoder(z.device)
        ctx.tape = _TapeSlot(vae, "_dtape_slots")
        slot = ctx.tape.slot
        img = nctx.decode_train(z, precision=vae._precision(), apply_scale_shift=apply_scale_shift, slot=slot)
        ctx.vae, ctx.nctx, ctx.slot = vae, nctx, slot
        ctx.names = [n for n, _ in vae.decoder.named_parameters()]
        ctx.z_dtype = z.dtype
        return img

    @staticmethod
    def backward(ctx, g_img):
        params = dict(ctx.vae.decoder.named_parameters())
        grads = {n: torch.empty_like(params[n], dtype=torch.float32) for n in ctx.names}
        gz = ctx.nctx.decoder_backward(g_img, grads, want_latent_grad=ctx.needs_input_grad[2], slot=ctx.slot)
        ctx.tape.release()
        return (None, None, None if gz is None else gz.to(ctx.z_dtype)) + tuple(grads[n].to(params[n].dtype) for n in ctx.names)


class AutoencoderKLOutput(SimpleNamespace):
    pass


class AutoencoderKL(nn.Module):
    """B200-native stand-in for ``diffusers.models.AutoencoderKL`` on the encode path."""

    def __init__(self, in_channels=3, out_channels=3, down_block_types=None, up_block_types=None,
                 block_out_channels=(128, 256, 512, 512), layers_per_block=2, act_fn="silu", latent_channels=16,
                 norm_num_groups=32, sample_size=1024, scaling_factor=0.3611, shift_factor=0.1159,
                 use_quant_conv=False, use_post_quant_conv=False, force_upcast=True, mid_block_add_attention=True,
                 latents_mean=None, latents_std=None, **unused):
        super().__init__()
        nblk = len(block_out_channels)
        if act_fn != "silu":
            raise ValueError("only act_fn='silu' is implemented (FLUX VAE)")
        if use_quant_conv or use_post_quant_conv:
            raise ValueError("use_quant_conv / use_post_quant_conv = True are not implemented "
                             "(the FLUX VAE of diffusers_vae_loader.py:102-134 has neither)")
        if down_block_types is not None and any(t != "DownEncoderBlock2D" for t in down_block_types):
            raise ValueError("only DownEncoderBlock2D encoder blocks are implemented")
        self.config = _Config(
            in_channels=in_channels, out_channels=out_channels,
            down_block_types=list(down_block_types or ["DownEncoderBlock2D"] * nblk),
            up_block_types=list(up_block_types or ["UpDecoderBlock2D"] * nblk),
            block_out_channels=list(block_out_channels), layers_per_block=layers_per_block, act_fn=act_fn,
            latent_channels=latent_channels, norm_num_groups=norm_num_groups, sample_size=sample_size,
            scaling_factor=scaling_factor, shift_factor=shift_factor, use_quant_conv=use_quant_conv,
            use_post_quant_conv=use_post_quant_conv, force_upcast=force_upcast,
            mid_block_add_attention=mid_block_add_attention, latents_mean=latents_mean, latents_std=latents_std)
        self.encoder = _encoder_params(self.config)
        # The decoder half (49.5 M parameters) is materialised on demand: by a checkpoint that carries
        # ``decoder.*`` keys, by ``decode()`` / ``forward()``, or by ``enable_decoder()``.
        self.decoder = None
        self._native_dec_key = None
        # "bf16" (tcgen05 path) or "fp32" (verification mode); VT_B200_PRECISION overrides the default
        self.precision = os.environ.get("VT_B200_PRECISION", "bf16")
        self.micro_batch = 0
        self.single_lane = False  # True: micro-batches back to back on one stream (per-kernel timing runs)
        self._native_key = None

    # ------------------------------------------------------------------ state dict
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = {}
        dropped = []
        if any(k.startswith("decoder.") for k in state_dict):
            self.enable_decoder()
        for k, v in state_dict.items():
            if k.startswith(("quant_conv.", "post_quant_conv.")):
                dropped.append(k)  # FLUX has no quant convs (use_quant_conv=False)
                continue
            parts = k.split(".")
            if "attentions" in parts and len(parts) >= 2 and parts[-2] in _LEGACY_ATTN:
                parts[-2:-1] = _LEGACY_ATTN[parts[-2]].split(".")
                k = ".".join(parts)
                if v.dim() == 4:  # legacy 1x1-conv attention projections
                    v = v.reshape(v.shape[0], v.shape[1])
            sd[k] = v
        result = super().load_state_dict(sd, strict=strict, assign=assign)
        self._native_key = None
        self._native_dec_key = None
        # quant_conv.* / post_quant_conv.* tensors have no home in this architecture: report them as the
        # reference's diffusers model would (diffusers_vae_loader.py:44-48 prints the unexpected keys)
        if dropped:
            if strict:
                raise RuntimeError(f"Unexpected key(s) in state_dict: {dropped}")
            result.unexpected_keys.extend(dropped)
        return result

    def enable_decoder(self):
        """Materialise the decoder parameters (PyTorch default init) next to the encoder's."""
        if self.decoder is None:
            ref = next(self.encoder.parameters())
            self.decoder = _decoder_params(self.config).to(device=ref.device, dtype=ref.dtype)
            for p in self.decoder.parameters():
                p.requires_grad_(ref.requires_grad)
        return self

    # ------------------------------------------------------------------ native mirror
    def _sync_native(self, device) -> "_native.Context":
        ctx = _native.get_context(device)
        params = list(self.encoder.named_parameters())
        key = (id(ctx), tuple((p.data_ptr(), p._version) for _, p in params))
        # the native context is process-wide per device and holds ONE encoder weight set: re-upload when these
        # parameters changed OR when another AutoencoderKL instance loaded its weights into the context since
        if key != self._native_key or getattr(ctx, "_enc_owner", None) is not self:
            ctx.configure_encoder(vars(self.config))
            ctx.load_encoder({n: p for n, p in params})
            self._native_key = key
            ctx._enc_owner = self
            ctx._dec_owner = None   # configure_encoder resets the shared configuration the decoder was built on
        return ctx

    def _precision(self) -> int:
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")
        return _native.PREC_FP32 if self.precision == "fp32" else _native.PREC_BF16

    def _device_of(self, x: torch.Tensor):
        if x.device.type != "cuda":
            raise _native.NativeError(
                "AutoencoderKL.encode needs CUDA tensors on a B200: the encode path has no CPU fallback")
        return x.device

    # ------------------------------------------------------------------ API
    def _take_tape_slot(self, pool="_tape_slots") -> int:
        used = self.__dict__.setdefault(pool, set())
        for s in range(_native.MAX_TAPES):
            if s not in used:
                used.add(s)
                return s
        raise RuntimeError(f"more than {_native.MAX_TAPES} encoder forwards are waiting for their backward")

    def _free_tape_slot(self, slot: int, pool="_tape_slots"):
        self.__dict__.setdefault(pool, set()).discard(slot)

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """``vae.encode(x).latent_dist`` (diffusers_vae_loader.py:73, :79).  In ``train()`` mode, with autograd enabled
        and trainable encoder parameters, the posterior carries a graph: the native training forward keeps its activations and the
        native backward produces every parameter gradient (the reference fine-tunes the VAE this way,
        train_full.py:201-256)."""
        self._device_of(x)
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.encoder.parameters()):
            params = [p for _, p in self.encoder.named_parameters()]
            mean, logvar = _EncodeTrainFn.apply(self, x, *params)
        else:
            with torch.no_grad():
                ctx = self._sync_native(x.device)
                _, mean, logvar = ctx.encode(x, precision=self._precision(), sample=False, apply_scale_shift=False,
                                             want_moments=True, micro_batch=self.micro_batch,
                                             single_lane=self.single_lane)
        dist = DiagonalGaussianDistribution(mean, logvar)
        if not return_dict:
            return (dist,)
        return AutoencoderKLOutput(latent_dist=dist)

    @torch.no_grad()
    def encode_latent(self, x: torch.Tensor, sample: bool = False, apply_scale_shift: bool = True, seed: int = 0,
                      noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused ``latent_dist.mode()|sample() * scaling_factor + shift_factor`` (one kernel epilogue);
        what ``DiffusersVAEWrapper.encode`` (diffusers_vae_loader.py:78-86) computes."""
        ctx = self._sync_native(self._device_of(x))
        return ctx.encode(x, precision=self._precision(), sample=sample, apply_scale_shift=apply_scale_shift,
                          seed=seed, noise=noise, micro_batch=self.micro_batch, single_lane=self.single_lane)

    def _sync_native_decoder(self, device) -> "_native.Context":
        ctx = self._sync_native(device)  # the decoder shares the encoder's configuration
        self.enable_decoder()
        params = list(self.decoder.named_parameters())
        key = (self._native_key, tuple((p.data_ptr(), p._version) for _, p in params))
        if key != self._native_dec_key or getattr(ctx, "_dec_owner", None) is not self:
            ctx.load_decoder({n: p for n, p in params})
            self._native_dec_key = key
            ctx._dec_owner = self
        return ctx

    def decode(self, z: torch.Tensor, return_dict: bool = True, apply_scale_shift: bool = False):
        """``vae.decode(z).sample`` (diffusers_vae_loader.py:75, :94); ``apply_scale_shift`` fuses the
        ``(z - shift_factor) / scaling_factor`` of ``DiffusersVAEWrapper.decode`` (:88-93) into the first kernel.
        In ``train()`` mode with autograd enabled the image carries a graph: gradients flow into the decoder
        parameters and back into ``z`` (train_vae.py:124-186)."""
        self._device_of(z)
        self.enable_decoder()
        if self.training and torch.is_grad_enabled() and (
                z.requires_grad or any(p.requires_grad for p in self.decoder.parameters())):
            params = [p for _, p in self.decoder.named_parameters()]
            img = _DecodeTrainFn.apply(self, bool(apply_scale_shift), z, *params)
        else:
            with torch.no_grad():
                ctx = self._sync_native_decoder(z.device)
                img = ctx.decode(z, precision=self._precision(), apply_scale_shift=apply_scale_shift,
                                 micro_batch=self.micro_batch)
        if not return_dict:
            return (img,)
        return DecoderOutput(sample=img)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None):
        """diffusers ``AutoencoderKL.forward``: decode(posterior.sample() or .mode())."""
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        return self.decode(z, return_dict=return_dict)

    def save_pretrained(self, path):
        """``config.json`` + ``diffusion_pytorch_model.safetensors`` in the diffusers layout (what the reference's
        training scripts call on the fine-tuned VAE, train_full.py:352); ``from_pretrained`` reads it back."""
        import json

        from safetensors.torch import save_file

        os.makedirs(path, exist_ok=True)
        cfg = {k: v for k, v in vars(self.config).items() if not k.startswith("_")}
        cfg["_class_name"] = "AutoencoderKL"
        with open(os.path.join(path, "config.json"), "w", encoding="utf-8") as f:
            json.dump(cfg, f, indent=2)
        save_file({k: v.detach().contiguous().cpu() for k, v in self.state_dict().items()},
                  os.path.join(path, "diffusion_pytorch_model.safetensors"))

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kw):
        """Local directory with ``config.json`` + ``diffusion_pytorch_model.safetensors`` (no network)."""
        import json

        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, "config.json"), "r", encoding="utf-8") as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        vae = cls(**cfg)
        st = os.path.join(root, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file

            vae.load_state_dict(load_file(st), strict=False)
        else:
            vae.load_state_dict(torch.load(os.path.join(root, "diffusion_pytorch_model.bin"), map_location="cpu"),
                                strict=False)
        return vae
