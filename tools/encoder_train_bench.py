"""one encoder fine-tuning step natively -- training forward (activations kept on a tape) + backward
(all 106 parameter gradients) -- timed on the device, with the per-class split of the backward.

    python tools/encoder_train_bench.py [--batch 4] [--res 512] [--steps 5] [--precision bf16]
Prints one JSON line.  Algorithmic work: forward F = 4.3329e12 p + 0.54976e12 p^2 FLOP per image (p = pixels / 1024^2,
SURVEY 8d); the backward is 2 F for the contractions that have both a data and a weight gradient (conv_in has only the
weight gradient) -- reported against 3 F per image for the step.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import _native  # noqa: E402
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    torch.manual_seed(0)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config()).cuda().train()
    vae.precision = a.precision
    x = torch.rand(a.batch, 3, a.res, a.res, device="cuda") * 2 - 1
    ctx = vae._sync_native(x.device)
    names = [n for n, _ in vae.encoder.named_parameters()]
    grads = {n: torch.empty_like(p, dtype=torch.float32) for n, p in vae.encoder.named_parameters()}
    prec = vae._precision()
    gm = torch.randn(a.batch, 16, a.res // 8, a.res // 8, device="cuda")
    gl = torch.randn_like(gm) * 0.1

    def fwd():
        return ctx.encode_train(x, precision=prec, slot=0)

    def bwd():
        ctx.encoder_backward(gm, gl, grads, slot=0)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for _ in range(a.warmup):
        fwd(); bwd()
    t_f = timed(fwd, a.steps)
    t_b = timed(bwd, a.steps)
    t_step = timed(lambda: (fwd(), bwd()), a.steps)
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    fwd()
    pf = ctx.profile_read(reset=True)
    bwd()
    pb = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    p = (a.res * a.res) / float(1024 * 1024)
    F = (4.3329e12 * p + 0.54976e12 * p * p) * a.batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1404.5))
    out = {
        "workload": f"encoder fine-tuning step, {a.batch} x {a.res}^2, {a.precision} mode (train_full.py:201-256)",
        "forward_ms": t_f, "backward_ms": t_b, "step_ms": t_step, "images_per_s": a.batch / t_step * 1e3,
        "algorithmic_tflop": {"forward": F / 1e12, "backward": 2 * F / 1e12},
        "achieved_tflops": {"forward": F / t_f / 1e9, "backward": 2 * F / t_b / 1e9, "step": 3 * F / t_step / 1e9},
        "frac_of_sustained_bf16_peak": {"forward": F / t_f / 1e9 / peak, "backward": 2 * F / t_b / 1e9 / peak,
                                        "step": 3 * F / t_step / 1e9 / peak},
        "backward_ms_by_class": {k: round(v["ms"], 3) for k, v in pb.items() if v["launches"]},
        "backward_launches_by_class": {k: int(v["launches"]) for k, v in pb.items() if v["launches"]},
        "forward_ms_by_class": {k: round(v["ms"], 3) for k, v in pf.items() if v["launches"]},
        "peak_tflops": peak,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
