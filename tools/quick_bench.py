"""Quick per-kernel-class timing of the encode+tag path with random weights (development aid)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from vae_tagger_b200 import _native as N  # noqa: E402
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
    dec = M.create_attention_decoder(16, R // 8, R // 8, 1000, attention_config={}).cuda().eval()
    x = torch.rand(B, 3, R, R, device="cuda") * 2 - 1
    ctx = N.get_context(0)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        lat = wrap.encode(x)
        out = dec.tag(lat)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"iter {it}: {dt * 1e3:.1f} ms  -> {B / dt:.2f} img/s  latent mean {lat.mean().item():.4f} std {lat.std().item():.4f}",
              flush=True)
    # lanes on/off comparison straight through the native context
    vae = wrap.vae
    nctx = vae._sync_native(x.device)
    for label, kw in (("two lanes mb=default", dict()), ("single lane mb=default", dict(single_lane=True)),
                      ("single lane mb=8", dict(single_lane=True, micro_batch=8)), ("two lanes mb=2", dict(micro_batch=2))):
        for _ in range(2):
            nctx.encode(x, **kw)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            nctx.encode(x, **kw)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 3
        print(f"encode only, {label:24s}: {dt * 1e3:7.2f} ms -> {B / dt:7.2f} img/s", flush=True)
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    lat = nctx.encode(x, single_lane=True)
    out = dec.tag(lat)
    prof = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    for k, v in prof.items():
        if v["launches"]:
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else 0
            gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else 0
            print(f"{k:16s} launches {int(v['launches']):5d}  {v['ms']:9.3f} ms  {tf:8.1f} TFLOP/s  {gb:8.1f} GB/s")
    print(json.dumps(prof))


if __name__ == "__main__":
    main()
