"""fp32 CPU restatement of the FLUX ``AutoencoderKL`` *decoder* path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned by the reference**, like
``oracle/encoder.py``: the arithmetic lives in the un-vendored ``diffusers`` package
(``requirements.txt:3``).  Restated from the published algorithm of diffusers'
``AutoencoderKL.decode`` -> ``Decoder.forward`` (``conv_in``, ``UNetMidBlock2D``, ``UpDecoderBlock2D`` x4
with ``layers_per_block + 1`` ``ResnetBlock2D`` each and ``Upsample2D`` = nearest 2x + conv3x3 on all but
the last block, ``conv_norm_out`` + SiLU + ``conv_out``), anchored on the reference's call sites:

  * ``vae.decode(z).sample`` ............................ diffusers_vae_loader.py:75, :94
  * ``(z - shift_factor) / scaling_factor`` first ....... diffusers_vae_loader.py:88-93
  * ``forward(x)`` = decode(latent_dist.sample()) ........ diffusers_vae_loader.py:72-76

Known answers pinned in tests: 49 545 475 decoder parameters (83 819 683 with the encoder),
138 tensors, output shape [B,3,8h,8w].
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder import FLUX_VAE_CONFIG, GN_EPS, OracleMidBlock, OracleResnetBlock2D


class OracleUpsample2D(nn.Module):
    """diffusers Upsample2D(use_conv=True): nearest-neighbour 2x, then conv3x3 pad 1."""

    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class OracleUpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList(
            [OracleResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([OracleUpsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class OracleDecoder(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(FLUX_VAE_CONFIG, **(cfg or {}))
        chans = list(reversed(cfg["block_out_channels"]))
        g = cfg["norm_num_groups"]
        self.conv_in = nn.Conv2d(cfg["latent_channels"], chans[0], 3, 1, 1)
        self.mid_block = OracleMidBlock(chans[0], g, cfg["mid_block_add_attention"])
        blocks, cin = [], chans[0]
        for i, cout in enumerate(chans):
            blocks.append(OracleUpDecoderBlock2D(cin, cout, cfg["layers_per_block"] + 1, g, i < len(chans) - 1))
            cin = cout
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(g, chans[-1], eps=GN_EPS, affine=True)
        self.conv_out = nn.Conv2d(chans[-1], cfg["out_channels"], 3, 1, 1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid_block(h)
        for blk in self.up_blocks:
            h = blk(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


def oracle_wrapper_decode(decoder: OracleDecoder, z: torch.Tensor, scaling_factor=0.3611, shift_factor=0.1159):
    """``DiffusersVAEWrapper.decode`` (diffusers_vae_loader.py:88-94): un-shift, un-scale, decode."""
    if shift_factor is not None:
        z = z - shift_factor
    if scaling_factor is not None:
        z = z / scaling_factor
    return decoder(z)


def make_oracle_decoder(seed: int = 0, cfg=None) -> OracleDecoder:
    """Random-init (PyTorch default init) oracle decoder under ``torch.manual_seed(seed)``."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    dec = OracleDecoder(cfg).eval()
    torch.random.set_rng_state(state)
    for p in dec.parameters():
        p.requires_grad_(False)
    return dec
