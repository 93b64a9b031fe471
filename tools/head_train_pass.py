"""One native head training step (B images, 128x128 latent, T tags) for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/head_train_pass.py 32"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import _native  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = 1000
ctx = _native.get_context(0)
torch.manual_seed(0)
dec = M.create_attention_decoder(16, 128, 128, T, attention_config={})
ctx.configure_head(_native.HEAD_ATTENTION, latent_channels=16, num_classes=T)
layout = ctx.head_param_layout()
sd = dec.state_dict()
flat = torch.cat([sd[n].reshape(-1) for n, _, _ in layout]).cuda()
grads = torch.zeros_like(flat)
m, v = torch.zeros_like(flat), torch.zeros_like(flat)
lat = torch.randn(B, 16, 128, 128, device="cuda") * 0.36 + 0.12
tgt = (torch.rand(B, T, device="cuda") < 0.1).float()
for it in range(2):
    ctx.head_train_step(lat, tgt, flat, grads, seed=it)
    ctx.adamw_step(flat, grads, m, v, lr=1e-3, step=it + 1, max_norm=1.0)
torch.cuda.synchronize()
print("ok")
