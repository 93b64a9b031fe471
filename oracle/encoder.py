"""fp32 CPU restatement of the FLUX ``AutoencoderKL`` *encoder* path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned by the
reference**: the arithmetic restated here lives in the third-party ``diffusers``
package (``requirements.txt:3`` ``diffusers>=0.21.0``; config dict declares
``_diffusers_version 0.30.0.dev0``, ``diffusers_vae_loader.py:105``), which is
absent from ``/root/reference`` and from this image.  The restatement follows the
published algorithm of diffusers' ``AutoencoderKL.encode`` ->
``Encoder.forward`` / ``DownEncoderBlock2D`` / ``ResnetBlock2D`` /
``Downsample2D`` / ``UNetMidBlock2D`` / ``Attention`` (AttnProcessor2_0) /
``DiagonalGaussianDistribution`` and is anchored on the
reference's own call sites:

  * construction + config keys ......... diffusers_vae_loader.py:7-35, :102-134
  * ``vae.encode(x).latent_dist`` ....... diffusers_vae_loader.py:73, :79
  * ``.mode()`` then ``* scaling_factor`` then ``+ shift_factor``
                                          diffusers_vae_loader.py:80-84
  * ``.sample()`` ....................... diffusers_vae_loader.py:74

Module / parameter names reproduce the diffusers state-dict keys (SURVEY.md
Appendix B) so FLUX ``diffusion_pytorch_model.safetensors`` files load.
Known answers pinned in tests: 34 274 208 parameters, 106 tensors.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

FLUX_VAE_CONFIG = {
    "in_channels": 3,
    "out_channels": 3,
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

GN_EPS = 1e-6


class OracleResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D (temb=None, dropout 0, output_scale_factor 1)."""

    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=GN_EPS, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=GN_EPS, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, 1, 0) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class OracleDownsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=0): pad right/bottom by 1, conv3x3 s2."""

    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, 2, 0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0.0))


class OracleDownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList(
            [OracleResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)]
        )
        self.downsamplers = nn.ModuleList([OracleDownsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class OracleAttention(nn.Module):
    """diffusers Attention(heads=1, dim_head=C, bias=True, residual_connection=True,
    norm_num_groups=groups) with AttnProcessor2_0 on a [B,C,H,W] input."""

    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=GN_EPS, affine=True)
        self.to_q = nn.Linear(ch, ch, bias=True)
        self.to_k = nn.Linear(ch, ch, bias=True)
        self.to_v = nn.Linear(ch, ch, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch, bias=True), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)  # [B,N,C]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(c)
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class OracleMidBlock(nn.Module):
    def __init__(self, ch, groups, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([OracleResnetBlock2D(ch, ch, groups) for _ in range(2)])
        self.attentions = nn.ModuleList([OracleAttention(ch, groups)]) if add_attention else None

    def forward(self, x):
        x = self.resnets[0](x)
        if self.attentions is not None:
            x = self.attentions[0](x)
        return self.resnets[1](x)


class OracleEncoder(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(FLUX_VAE_CONFIG, **(cfg or {}))
        chans = list(cfg["block_out_channels"])
        g = cfg["norm_num_groups"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], chans[0], 3, 1, 1)
        blocks, cin = [], chans[0]
        for i, cout in enumerate(chans):
            blocks.append(
                OracleDownEncoderBlock2D(cin, cout, cfg["layers_per_block"], g, i < len(chans) - 1)
            )
            cin = cout
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = OracleMidBlock(chans[-1], g, cfg["mid_block_add_attention"])
        self.conv_norm_out = nn.GroupNorm(g, chans[-1], eps=GN_EPS, affine=True)
        self.conv_out = nn.Conv2d(chans[-1], 2 * cfg["latent_channels"], 3, 1, 1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            h = blk(h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class OracleDiagonalGaussian:
    """diffusers DiagonalGaussianDistribution."""

    def __init__(self, moments: torch.Tensor):
        self.mean, logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def kl(self):
        return 0.5 * torch.sum(self.mean ** 2 + self.var - 1.0 - self.logvar, dim=[1, 2, 3])


class OracleAutoencoderKL(nn.Module):
    """Encoder half of diffusers AutoencoderKL (no quant_conv: ``use_quant_conv=False``)."""

    def __init__(self, cfg=None):
        super().__init__()
        full = dict(FLUX_VAE_CONFIG, **(cfg or {}))
        self.config = SimpleNamespace(**full)
        self.encoder = OracleEncoder(full)

    def encode(self, x):
        return SimpleNamespace(latent_dist=OracleDiagonalGaussian(self.encoder(x)))


def oracle_wrapper_encode(vae: OracleAutoencoderKL, x: torch.Tensor) -> torch.Tensor:
    """``DiffusersVAEWrapper.encode`` (diffusers_vae_loader.py:78-86): mode()*scale + shift."""
    latent = vae.encode(x).latent_dist.mode()
    if hasattr(vae.config, "scaling_factor"):
        latent = latent * vae.config.scaling_factor
    if hasattr(vae.config, "shift_factor"):
        latent = latent + vae.config.shift_factor
    return latent


def make_oracle_vae(seed: int = 0, cfg=None) -> OracleAutoencoderKL:
    """Random-init (PyTorch default init) oracle VAE under ``torch.manual_seed(seed)``."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    vae = OracleAutoencoderKL(cfg).eval()
    torch.random.set_rng_state(gen_state)
    for p in vae.parameters():
        p.requires_grad_(False)
    return vae


def synthetic_images(n: int, h: int, w: int, seed: int = 1234) -> torch.Tensor:
    """image i ~ U[-1,1) from Generator().manual_seed(seed+i), fp32 NCHW."""
    out = torch.empty(n, 3, h, w)
    for i in range(n):
        g = torch.Generator().manual_seed(seed + i)
        out[i] = torch.rand(3, h, w, generator=g) * 2.0 - 1.0
    return out


def structured_images(n: int, h: int, w: int, seed: int = 4321, return_params: bool = False):
    """Varied synthetic "photo-like" images in [-1,1] (fp32 NCHW): a smooth random colour field (bicubic
    up-sampling of a small random grid) times a random contrast, plus a random colour offset, plus fine noise
    of a random amplitude.  Uniform-noise images (``synthetic_images``) all give nearly the same latent
    statistics, so the tag head sees nearly the same input for every image; these spread the latents, and with
    them the logits, so that a per-image tag-set comparison is a comparison of different cases.
    ``return_params``: also the generating parameters per image [n, 6] = (cells/8, contrast, offset r, g, b,
    noise amplitude) -- what ``make_tagset_golden.py`` derives its synthetic training labels from."""
    out = torch.empty(n, 3, h, w)
    params = torch.empty(n, 6)
    for i in range(n):
        g = torch.Generator().manual_seed(seed + i)
        cells = int(torch.randint(2, 9, (1,), generator=g))
        grid = torch.rand(1, 3, cells, cells, generator=g) * 2.0 - 1.0
        field = F.interpolate(grid, size=(h, w), mode="bicubic", align_corners=False)[0]
        contrast = 0.2 + 0.8 * float(torch.rand(1, generator=g))
        offset = (torch.rand(3, 1, 1, generator=g) * 2.0 - 1.0) * 0.5
        noise_amp = 0.02 + 0.4 * float(torch.rand(1, generator=g)) ** 2
        noise = (torch.rand(3, h, w, generator=g) * 2.0 - 1.0) * noise_amp
        out[i] = (field * contrast + offset + noise).clamp_(-1.0, 1.0)
        params[i] = torch.tensor([cells / 8.0, contrast, *offset.flatten().tolist(), noise_amp])
    return (out, params) if return_params else out
