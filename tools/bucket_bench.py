"""BASELINE config 3: aspect-ratio bucketed inference, every reachable bucket (512..1024 step 64).

For each bucket (W, H): one batch of ``--mpx`` megapixels (at least 2 images) of synthetic images through
encode + tag, CUDA-event timed after warm-up; reports images/s, algorithmic TFLOP/s (SURVEY 8d:
4.3329e12*p + 0.54976e12*p^2 FLOP per image, p = W*H/1024^2) and the fraction of the measured bf16 peak.
Writes one JSON line per bucket and a summary line (mixed-bucket aggregate = total images / total time).

    python tools/bucket_bench.py [--mpx 16] [--out gpurun_out/buckets.jsonl]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402


def flops(w, h):
    p = w * h / (1024.0 * 1024.0)
    return 4.3329e12 * p + 0.54976e12 * p * p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mpx", type=float, default=16.0)
    ap.add_argument("--out", default="gpurun_out/buckets.jsonl")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    peaks = {"burst": 1703.9, "sustained": 1404.5}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            d = json.load(f)
        peaks["burst"] = d.get("bf16_tflops", d.get("bf16_tflops_burst", peaks["burst"]))
        peaks["sustained"] = d.get("bf16_tflops_sustained", peaks["sustained"])
    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
    arb = M.AspectRatioBucketing()
    reach = sorted({arb.bucket_for_size(w, h) for (w, h) in arb.buckets})
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    heads = {}
    tot_img, tot_ms, tot_flop = 0, 0.0, 0.0
    with open(a.out, "w") as f:
        for (w, h) in reach:
            n = max(2, int(a.mpx * 1e6 / (w * h)))
            key = (h // 8, w // 8)
            if key not in heads:
                heads[key] = M.create_attention_decoder(16, h // 8, w // 8, 1000, attention_config={}).cuda().eval()
            dec = heads[key]
            x = torch.rand(n, 3, h, w, device="cuda") * 2 - 1
            for _ in range(2):
                out = dec.tag(wrap.encode(x))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                out = dec.tag(wrap.encode(x))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            tf = flops(w, h) * n / (ms * 1e-3) / 1e12
            line = {"bucket": [w, h], "images": n, "ms": ms, "images_per_s": n / (ms * 1e-3), "tflops": tf,
                    "frac_sustained": tf / peaks["sustained"], "frac_burst": tf / peaks["burst"]}
            f.write(json.dumps(line) + "\n")
            tot_img += n
            tot_ms += ms
            tot_flop += flops(w, h) * n
            del x, out
        summ = {"summary": True, "buckets": len(reach), "images": tot_img, "ms": tot_ms,
                "images_per_s": tot_img / (tot_ms * 1e-3), "tflops": tot_flop / (tot_ms * 1e-3) / 1e12,
                "frac_sustained": tot_flop / (tot_ms * 1e-3) / 1e12 / peaks["sustained"],
                "frac_burst": tot_flop / (tot_ms * 1e-3) / 1e12 / peaks["burst"], "peaks": peaks}
        f.write(json.dumps(summ) + "\n")
    print(json.dumps(summ))


if __name__ == "__main__":
    main()
