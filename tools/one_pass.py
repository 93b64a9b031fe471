"""One encode+tag pass with random weights (profiling target: ncu launch lists / full captures).

    python tools/one_pass.py [batch] [resolution] [passes]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
P = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(0)
wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
dec = M.create_attention_decoder(16, R // 8, R // 8, 1000, attention_config={}).cuda().eval()
x = torch.rand(B, 3, R, R, device="cuda") * 2 - 1
for _ in range(P):
    lat = wrap.encode(x)
    out = dec.tag(lat)
torch.cuda.synchronize()
print("ok", lat.mean().item(), int(out["count"].sum().item()))
