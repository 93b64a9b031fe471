"""Development aid: the encode / tag / host paths on image sizes that are multiples of 8 only, uint8 vs float input,
tiny latents.  (Parity against the oracle for these shapes is in tests/test_gpu_encoder.py.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402
from vae_tagger_b200.infer_full import encode_and_tag  # noqa: E402

torch.manual_seed(0)
wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


for (B, H, W) in [(2, 72, 88), (1, 8, 8), (3, 40, 8), (1, 264, 1000), (37, 64, 64), (1, 9, 15), (2, 301, 203), (1, 500, 500)]:
    g = torch.Generator().manual_seed(H * W)
    xu = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    xf = ((xu.float() / 255 - 0.5) / 0.5).permute(0, 3, 1, 2).contiguous()
    dec = M.create_attention_decoder(16, H // 8, W // 8, 11, attention_config={}).cuda().eval()
    out = {}
    for prec in ("fp32", "bf16"):
        wrap.vae.precision = prec
        a = wrap.encode(xu.cuda())
        b = wrap.encode(xf.cuda())
        t = encode_and_tag(wrap, dec, xu.cuda())
        ctx = wrap.vae._sync_native(torch.device("cuda", 0))
        h = ctx.infer_host(xu.pin_memory(), precision=0 if prec == "bf16" else 1, want_latent=True)
        out[prec] = (f"u8-vs-f32 {rel(a, b):.1e}", f"fused-vs-encode {rel(t['latent'], a):.1e}",
                     f"host-vs-device conf {(h['conf'] - t['conf'].cpu()).abs().max().item():.1e}",
                     f"finite {bool(torch.isfinite(t['conf']).all())}")
    print((B, H, W), out, flush=True)
