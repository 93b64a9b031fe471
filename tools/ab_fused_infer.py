"""A/B on one box: fused vt_infer (head per micro-batch behind its encoder) vs encode() then tag(),
batch 32 at 1024^2, alternating blocks of steps (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402
from vae_tagger_b200.infer_full import encode_and_tag  # noqa: E402

B, R, T = 32, 1024, 1000
torch.manual_seed(0)
wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
dec = M.create_attention_decoder(16, R // 8, R // 8, T, attention_config={}).cuda().eval()
x = torch.rand(B, 3, R, R, device="cuda") * 2 - 1


def two_calls():
    return dec.tag(wrap.encode(x), threshold=0.5)


def fused():
    return encode_and_tag(wrap, dec, x, threshold=0.5)


def timed(fn, steps=4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for _ in range(3):
    two_calls(); fused()
res = {"two_calls": [], "fused": []}
for _ in range(4):
    res["two_calls"].append(timed(two_calls))
    res["fused"].append(timed(fused))
for k, v in res.items():
    print(k, [round(t, 2) for t in v], "ms/step; best", round(B / min(v) * 1e3, 1), "img/s")
