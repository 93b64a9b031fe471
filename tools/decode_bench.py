"""Throughput of the native VAE decoder: images/s for ``DiffusersVAEWrapper.decode`` at
--res (latent res/8), batch B, bf16 tensor-core mode, with the algorithmic FLOP roofline (parity against the
oracle is the tests' job: tests/test_gpu_decoder.py).

    python tools/decode_bench.py [--res 1024] [--batch 8] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402


def decoder_flops(h, w, chans=(128, 256, 512, 512), layers=2, lc=16):
    """Algorithmic FLOPs (2*MAC) of one image through diffusers' Decoder at latent h x w."""
    rc = list(reversed(chans))
    conv = lambda hh, ww, ci, co, k=9: 2.0 * hh * ww * co * ci * k  # noqa: E731
    f = conv(h, w, lc, rc[0])
    c = rc[0]
    f += 4 * conv(h, w, c, c)                                   # two mid resnets
    n = h * w
    f += 4 * 2.0 * n * c * c + 2 * 2.0 * n * n * c              # q,k,v,out projections + QK^T + PV
    cin = c
    for i, co in enumerate(rc):
        for j in range(layers + 1):
            ci = cin if j == 0 else co
            f += conv(h, w, ci, co) + conv(h, w, co, co)
            if ci != co:
                f += conv(h, w, ci, co, 1)
        if i < len(rc) - 1:
            h, w = 2 * h, 2 * w
            f += conv(h, w, co, co)
        cin = co
    return f + conv(h, w, rc[-1], 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config()).enable_decoder())
    wrap = wrap.cuda().eval()
    h = a.res // 8
    z = torch.randn(a.batch, 16, h, h, device="cuda") * 0.36 + 0.12

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    ms = timed(lambda: wrap.decode(z))
    x = torch.rand(a.batch, 3, a.res, a.res, device="cuda") * 2 - 1
    ms_rt = timed(lambda: wrap(x))
    flops = decoder_flops(h, h) * a.batch
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "MEASURED_PEAKS.json")))
    peak = float(peaks["bf16_tflops_sustained"])
    print(json.dumps({
        "metric": "images/s VAE decode bf16", "workload": f"latent {h}x{h} -> {a.res}x{a.res}, batch {a.batch}",
        "value": round(a.batch / ms * 1e3, 2), "ms_per_step": round(ms, 3),
        "reconstruct_images_per_s": round(a.batch / ms_rt * 1e3, 2), "reconstruct_ms_per_step": round(ms_rt, 3),
        "roofline": {"bound": "tensor", "achieved": round(flops / ms / 1e9, 1), "peak": peak, "unit": "TFLOP/s",
                     "frac": round(flops / ms / 1e9 / peak, 4), "algorithmic_flop_per_image": flops / a.batch,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
    }))


if __name__ == "__main__":
    main()
