"""Throughput of the GPU preprocessing kernels against their HBM roofline, with Pillow on
the host cores beside it.  Workload: decoded 8-bit RGB photos of --src WxH, SmartResize'd into their
aspect-ratio bucket (LANCZOS), B images per step, sources resident in HBM (value) or uploaded from pinned
host memory inside the timed region (e2e).

    python tools/preprocess_bench.py [--src 3000x2000] [--batch 32] [--steps 10]
Algorithmic bytes per image = cropped source read once + destination written once (the uint8 intermediate
of the two-pass resize is extra traffic the roofline fraction pays for).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import _native  # noqa: E402
from vae_tagger_b200.modules import AspectRatioBucketing, SmartResize  # noqa: E402
from vae_tagger_b200.preprocess import gpu_smart_resize  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="3000x2000")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    sw, sh = (int(v) for v in a.src.split("x"))
    ctx = _native.get_context(0)
    tw, th = AspectRatioBucketing().bucket_for_size(sw, sh)
    box = _native.smart_crop_box(sw, sh, tw, th)
    rng = np.random.default_rng(0)
    host = torch.from_numpy(rng.integers(0, 256, (a.batch, sh, sw, 3), dtype=np.uint8)).pin_memory()
    dev = host.cuda()
    out = torch.empty(a.batch, th, tw, 3, dtype=torch.uint8, device="cuda")

    def step(src):  # one native call per bucket batch
        ctx.resize_u8_batch([src[i] for i in range(a.batch)], (tw, th), [box] * a.batch, _native.FILTER_LANCZOS, out=out)

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

    ms = timed(lambda: step(dev))
    staging = torch.empty_like(dev)

    def e2e():
        staging.copy_(host, non_blocking=True)
        step(staging)

    ms_e2e = timed(e2e)
    cw, ch = box[2] - box[0], box[3] - box[1]
    alg = a.batch * 3.0 * (cw * ch + tw * th)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "MEASURED_PEAKS.json")))
    peak = float(peaks.get("hbm_gbs", 6446.0))
    # host baseline: the reference's own transform (PIL) on a bounded sample, one thread per image is what
    # a DataLoader worker does; report per-core and x cores
    from PIL import Image

    sample = [Image.fromarray(host[i].numpy()) for i in range(min(4, a.batch))]
    sr = SmartResize(tw, th)
    t0 = time.perf_counter()
    for im in sample:
        sr(im)
    cpu_s = (time.perf_counter() - t0) / len(sample)
    print(json.dumps({
        "metric": "images/s SmartResize (crop + LANCZOS) uint8", "workload": f"{sw}x{sh} -> {tw}x{th}, batch {a.batch}",
        "value": round(a.batch / ms * 1e3, 1), "e2e": {"value": round(a.batch / ms_e2e * 1e3, 1),
                                                        "h2d_bytes_per_step": host.numel(), "d2h_bytes_per_step": 0},
        "ms_per_step": round(ms, 3), "gpu_launches": 2 * a.batch * a.steps,
        "roofline": {"bound": "hbm", "achieved": round(alg / ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(alg / ms / 1e6 / peak, 4), "algorithmic_bytes_per_image": alg / a.batch},
        "cpu_baseline": {"value": round(1.0 / cpu_s, 2), "unit": "images/s", "cores": 1, "kind": "reference",
                         "sample": f"{len(sample)} images through SmartResize (PIL {Image.__version__}) on one core"},
    }))


if __name__ == "__main__":
    main()
