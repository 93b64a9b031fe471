"""Drop-in for the reference's ``diffusers_vae_loader.py`` on the encode path.

Same public names, argument meaning and error behaviour
(``/root/reference/diffusers_vae_loader.py``): ``load_diffusers_vae_from_config`` (:7-53),
``load_diffusers_vae_from_pretrained`` (:55-65), ``DiffusersVAEWrapper`` (:67-94),
``create_vae_from_config_file`` (:96-100), ``get_diffusers_vae_config`` (:102-134) -- but the
``AutoencoderKL`` they build is :class:`vae_tagger_b200.autoencoder_kl.AutoencoderKL`, whose
encoder runs as hand-written sm_100a kernels.
"""
from __future__ import annotations

import json
import os

import torch

from .autoencoder_kl import AutoencoderKL

_FLUX_BLOCKS_DOWN = ["DownEncoderBlock2D"] * 4
_FLUX_BLOCKS_UP = ["UpDecoderBlock2D"] * 4

# the keys the reference forwards to AutoencoderKL(...) with their defaults (:9-34)
_CONFIG_DEFAULTS = {
    "in_channels": 3,
    "out_channels": 3,
    "down_block_types": _FLUX_BLOCKS_DOWN,
    "up_block_types": _FLUX_BLOCKS_UP,
    "block_out_channels": [128, 256, 512, 512],
    "layers_per_block": 2,
    "act_fn": "silu",
    "latent_channels": 16,
    "norm_num_groups": 32,
    "sample_size": 1024,
    "scaling_factor": 0.3611,
    "shift_factor": 0.1159,
    "use_quant_conv": False,
    "use_post_quant_conv": False,
    "force_upcast": True,
    "mid_block_add_attention": True,
}


def load_diffusers_vae_from_config(config_dict, model_path=None):
    """Build the VAE from a config mapping; load weights when ``model_path`` exists.

    As in the reference, a missing ``model_path`` silently leaves the random initialisation
    (:37) and weights load with ``strict=False``, reporting missing / unexpected keys (:44-49)."""
    kwargs = {k: config_dict.get(k, d) for k, d in _CONFIG_DEFAULTS.items()}
    vae = AutoencoderKL(**kwargs)
    if model_path and os.path.exists(model_path):
        print(f"loading pretrained VAE weights: {model_path}")
        if model_path.endswith(".safetensors"):
            from safetensors.torch import load_file as load_safetensors

            state_dict = load_safetensors(model_path)
        else:
            state_dict = torch.load(model_path, map_location="cpu")
        missing_keys, unexpected_keys = vae.load_state_dict(state_dict, strict=False)
        if missing_keys:
            print(f"missing keys: {missing_keys}")
        if unexpected_keys:
            print(f"unexpected keys: {unexpected_keys}")
        print("pretrained VAE weights loaded")
    return vae


def load_diffusers_vae_from_pretrained(model_name_or_path, subfolder=None):
    """Local-directory ``from_pretrained``; returns ``None`` on failure like the reference (:62-65)."""
    try:
        vae = AutoencoderKL.from_pretrained(model_name_or_path, subfolder=subfolder)
        print(f"loaded pretrained VAE from {model_name_or_path}")
        return vae
    except Exception as e:  # noqa: BLE001 - the reference swallows every failure here
        print(f"loading VAE from {model_name_or_path} failed: {e}")
        return None


class DiffusersVAEWrapper(torch.nn.Module):
    """``encode(x) = latent_dist.mode() * scaling_factor + shift_factor`` (:78-86)."""

    def __init__(self, vae_model):
        super().__init__()
        self.vae = vae_model

    def forward(self, x):
        # (:72-76) reconstruction through the native decoder
        posterior = self.vae.encode(x).latent_dist
        z = posterior.sample()
        reconstruction = self.vae.decode(z).sample
        return reconstruction, posterior

    def encode(self, x):
        if isinstance(self.vae, AutoencoderKL):
            # scale and shift are fused into the moments->latent kernel; the hasattr checks of the
            # reference (:81-84) are honoured by the config the native context was given
            return self.vae.encode_latent(x, sample=False, apply_scale_shift=True)
        latent = self.vae.encode(x).latent_dist.mode()
        if hasattr(self.vae.config, "scaling_factor"):
            latent = latent * self.vae.config.scaling_factor
        if hasattr(self.vae.config, "shift_factor"):
            latent = latent + self.vae.config.shift_factor
        return latent

    def decode(self, z):
        if isinstance(self.vae, AutoencoderKL):
            # un-shift / un-scale (:89-93) fused into the decoder's first kernel
            return self.vae.decode(z, apply_scale_shift=True).sample
        if hasattr(self.vae.config, "shift_factor"):
            z = z - self.vae.config.shift_factor
        if hasattr(self.vae.config, "scaling_factor"):
            z = z / self.vae.config.scaling_factor
        return self.vae.decode(z).sample


def create_vae_from_config_file(config_path, model_path=None):
    with open(config_path, "r", encoding="utf-8") as f:
        config = json.load(f)
    return DiffusersVAEWrapper(load_diffusers_vae_from_config(config, model_path))


def get_diffusers_vae_config():
    """The canonical FLUX VAE config dict (:102-134)."""
    cfg = {"_class_name": "AutoencoderKL", "_diffusers_version": "0.30.0.dev0"}
    cfg.update({k: (list(v) if isinstance(v, list) else v) for k, v in _CONFIG_DEFAULTS.items()})
    cfg.update({"latents_mean": None, "latents_std": None})
    return dict(sorted(cfg.items()))
