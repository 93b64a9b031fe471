"""Batch-invariance / run-to-run check of the 16-bit encoder path.

Encodes the same images (a) twice, (b) with different micro-batch splits, (c) in a different batch order and at a
different batch size, and reports for each comparison whether the latents are BIT-identical (and the max abs
difference if not).  One JSON line.  Usage: python tools/determinism_check.py [H W]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.encoder import make_oracle_vae, structured_images, synthetic_images  # noqa: E402
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    oracle = make_oracle_vae(0)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    vae.load_state_dict(oracle.state_dict(), strict=False)
    wrap = L.DiffusersVAEWrapper(vae).cuda().eval()
    x = torch.cat([synthetic_images(3, H, W), structured_images(3, H, W)]).cuda()
    out = {"H": H, "W": W}
    for prec in ("bf16", "fp32"):
        wrap.vae.precision = prec
        wrap.vae.micro_batch = 0
        a = wrap.encode(x).clone()

        def cmp(name, b):
            same = torch.equal(a, b)
            out[f"{prec}:{name}"] = "bit-identical" if same else f"max|d| {(a - b).abs().max().item():.3e}"

        cmp("rerun", wrap.encode(x).clone())
        for mb in (1, 2, 4, 6):
            wrap.vae.micro_batch = mb
            cmp(f"micro_batch={mb}", wrap.encode(x).clone())
        wrap.vae.micro_batch = 0
        perm = torch.tensor([4, 2, 0, 5, 1, 3], device=x.device)
        b = torch.empty_like(a)
        b[perm] = wrap.encode(x[perm].contiguous())
        cmp("permuted", b)
        b = torch.cat([wrap.encode(x[:1].contiguous()), wrap.encode(x[1:].contiguous())])
        cmp("split 1+5", b)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
