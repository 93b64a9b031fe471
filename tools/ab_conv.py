"""A/B timing of single conv kernels through the op entry points (development aid).
    VT_B200_LIB=path/to/lib.so python tools/ab_conv.py
Prints the CUDA-event time of the contraction launch only (profiler class igemm)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import _native as N  # noqa: E402

ctx = N.get_context(0)
torch.manual_seed(0)
cases = [("fused 128->128 @1024^2 x2", 2, 128, 1024, 128), ("fused 256->256 @512^2 x2", 2, 256, 512, 256),
         ("fused 512->512 @256^2 x2", 2, 512, 256, 512), ("fused 512->512 @128^2 x4", 4, 512, 128, 512)]
for name, n, c, r, co in cases:
    x = torch.randn(n, c, r, r, device="cuda")
    g = torch.ones(c, device="cuda"); b = torch.zeros(c, device="cuda")
    w = torch.randn(co, c, 3, 3, device="cuda") * 0.03
    bias = torch.zeros(co, device="cuda")
    res = torch.randn(n, co, r, r, device="cuda")
    for use_res in (False, True):
        ts = []
        for it in range(6):
            ctx.profile_enable(True); ctx.profile_read(reset=True)
            ctx.op_conv3_fused(x, g, b, w, bias, res if use_res else None, want_stats=True)
            p = ctx.profile_read(reset=True)
            ts.append(p["igemm_tcgen05"]["ms"])
        ctx.profile_enable(False)
        ts = sorted(ts[1:])
        fl = 2.0 * n * r * r * co * 9 * c
        print(f"{os.path.basename(N.lib_path()):24s} {name:28s} res={int(use_res)}  median {ts[len(ts)//2]*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us  {fl/ts[len(ts)//2]/1e9:7.0f} TF/s", flush=True)
    del x, w, res
