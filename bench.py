"""bench.py -- images/sec of the encode+tag hot path at 1024x1024 bf16 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path (FLUX VAE encoder -> latent mode*0.3611+0.1159 -> attention tag
head -> sigmoid / sort / threshold 0.5) over one batch of 32 synthetic 1024x1024 images per GPU
(BASELINE.json configs[1]; random-init weights, torch.manual_seed(0); images U[-1,1)).
`value` is the whole-job images/s with inputs resident in HBM; `e2e` is the same metric through
the C-ABI host call (pinned host images H2D, results D2H inside the timed region).  Rank r of N
processes its own batch (weak scaling, no collective on the data path).

`--impl reference` times the CPU restatement of the reference's path (oracle/, "port": diffusers is
not installable in this image) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec encode+tag at 1024^2 bf16"
UNIT = "images/s"
RES = 1024
BATCH = 32
NUM_TAGS = 1000
FLOP_PER_IMAGE = 4.8826e12  # encoder contractions per 1024^2 image


def flops_per_image(res: int) -> float:
    p = (res * res) / (1024.0 * 1024.0)
    return 4.3329e12 * p + 0.54976e12 * p * p


def make_config(B, R, world):
    return {"workload": f"configs[1]: {R}x{R} batch {B} per GPU bf16 encode+tag (infer_full.py path), "
                        f"FLUX VAE encoder random init + 8-head attention tagger, {NUM_TAGS} tags",
            "global_batch": B * world, "resolution": R, "parallelism": f"dp{world} (batch sharded, no collective)",
            "l2": "inputs larger than L2 (403 MB image batch, GB-scale activations)"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"burst": float(d.get("bf16_tflops", 1590.0)), "sustained": float(d.get("bf16_tflops_sustained", 1400.0)),
                "hbm": float(d.get("hbm_gbs", 6650.0)), "source": "MEASURED_PEAKS.json"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_ident):   # nvidia-smi -i accepts an index or a "GPU-<uuid>"
        self.gpu = gpu_ident
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        rows, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, pw = float(f[1]), float(f[3])
                mx.append(float(f[2]))
            except ValueError:
                continue
            rows.append((clk, pw))
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        # "under load" = samples drawing at least 60 % of the highest power seen: the sampler also catches the idle
        # moments around the timed region (barriers, its own start-up), where the clock sits at its maximum
        pmax = max((pw for _, pw in rows), default=0.0)
        load = [clk for clk, pw in rows if pw >= 0.6 * pmax]
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(load), "samples_total": len(rows), "power_w_max": pmax, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(budget_s: float, threads: int):
    """Oracle (CPU restatement of the reference path) timed on the host cores on a bounded sample.
    Returns (images/s scaled to 1024^2, sample description)."""
    import torch

    from oracle import encoder as OE
    from oracle import head as OH

    torch.set_num_threads(threads)
    vae = OE.make_oracle_vae(0)
    torch.manual_seed(0)
    # head weights in the reference layout (random init, shapes per SURVEY.md Appendix B)
    sd = {
        "spatial_attention.channel_att.0.weight": torch.randn(2, 16, 1, 1) * 0.2,
        "spatial_attention.channel_att.2.weight": torch.randn(16, 2, 1, 1) * 0.2,
        "spatial_attention.spatial_att.0.weight": torch.randn(1, 2, 7, 7) * 0.1,
        "feature_compress.0.weight": torch.randn(8, 16, 3, 3) * 0.1, "feature_compress.0.bias": torch.zeros(8),
        "feature_compress.1.weight": torch.ones(8), "feature_compress.1.bias": torch.zeros(8),
        "feature_compress.1.running_mean": torch.zeros(8), "feature_compress.1.running_var": torch.ones(8),
    }
    for k in ("q_proj", "k_proj", "v_proj", "out_proj"):
        sd[f"self_attention_post.{k}.weight"] = torch.randn(8, 8) * 0.3
        sd[f"self_attention_post.{k}.bias"] = torch.zeros(8)
    sd["self_attention_post.norm.weight"] = torch.ones(8); sd["self_attention_post.norm.bias"] = torch.zeros(8)
    dims = [512, 1024, 512, 256, NUM_TAGS]
    for i, (lin, ln) in enumerate(((0, 1), (4, 5), (8, 9), (12, None))):
        sd[f"classifier.{lin}.weight"] = torch.randn(dims[i + 1], dims[i]) / dims[i] ** 0.5
        sd[f"classifier.{lin}.bias"] = torch.zeros(dims[i + 1])
        if ln is not None:
            sd[f"classifier.{ln}.weight"] = torch.ones(dims[i + 1]); sd[f"classifier.{ln}.bias"] = torch.zeros(dims[i + 1])

    def one(res):
        x = OE.synthetic_images(1, res, res)
        t0 = time.perf_counter()
        with torch.no_grad():
            lat = OE.oracle_wrapper_encode(vae, x)
            conf, idx = OH.get_confidence(OH.attention_decoder_logits(sd, lat))
            OH.threshold_tags(conf[0], idx[0], 0.5)
        return time.perf_counter() - t0

    t256 = one(256)  # probe (also warms the thread pool)
    rate = flops_per_image(256) / t256
    res = 256
    for cand in (1024, 768, 512, 384):
        if flops_per_image(cand) / rate * 1.3 <= budget_s:
            res = cand
            break
    t = one(res) if res != 256 else t256
    ips_1024 = (1.0 / t) * (flops_per_image(res) / FLOP_PER_IMAGE)
    sample = (f"1 image {res}x{res} fp32 through oracle encoder + head ({t:.2f} s, {threads} threads)"
              + ("" if res == 1024 else f", scaled to 1024^2 by the FLOP ratio {flops_per_image(res) / FLOP_PER_IMAGE:.4f}"))
    return ips_1024, sample, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_step_budget = max(2.0, 150.0 / (steps + warm))
    vals, sample = [], ""
    t_wall0 = time.perf_counter()
    for i in range(warm + steps):
        ips, sample, _ = cpu_reference_rate(per_step_budget, threads)
        if i >= warm:
            vals.append(ips)
        if time.perf_counter() - t_wall0 > 240 and vals:
            break
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": warm, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(args.batch, args.resolution, max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from vae_tagger_b200 import _native
    from vae_tagger_b200 import diffusers_vae_loader as L
    from vae_tagger_b200 import modules as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200): the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, R = args.batch, args.resolution

    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).to(dev).eval()
    dec = M.create_attention_decoder(16, R // 8, R // 8, NUM_TAGS, attention_config={}).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host = (torch.rand(B, 3, R, R, generator=g) * 2 - 1).pin_memory()
    x = host.to(dev, non_blocking=True)
    ctx = _native.get_context(dev)

    from vae_tagger_b200.infer_full import encode_and_tag

    def step():  # the call infer_full.py makes per batch: encode + get_confidence + threshold count
        return encode_and_tag(wrap, dec, x, threshold=0.5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(3, args.warmup)):
        out = step()
    ctx.profile_read(reset=True)
    # nvidia-smi numbers the GPUs by PCI bus, CUDA by its own order (and CUDA_VISIBLE_DEVICES): name the device by UUID
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        gpu_ident = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    except Exception:  # noqa: BLE001 - older torch: fall back to the index
        gpu_ident = str(local)
    sampler = ClockSampler(gpu_ident)
    if rank == 0:
        sampler.start()
    ms = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    prof = ctx.profile_read(reset=True)
    launches = int(sum(v["launches"] for v in prof.values()))
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the C-ABI host call: pinned host images -> H2D -> encode -> tag -> D2H
    e2e_out = None

    def e2e_step():
        nonlocal e2e_out
        e2e_out = ctx.infer_host(host, threshold=0.5, out=e2e_out)

    _native.get_context(dev)  # same context: the head/encoder parameters are already resident
    e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns after the D2H copies have landed
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = t.item()
    e2e_value = world * B * e2e_steps / t_e2e
    h2d = host.numel() * host.element_size()
    d2h = B * NUM_TAGS * (4 + 8) + B * 4

    # ---- roofline of the dominant kernel (tcgen05 implicit GEMM), per-launch CUDA events
    # (micro-batches back to back on one stream here: with the two overlapping lanes of the timed
    # region the per-launch event intervals would include the other lane's kernels)
    wrap.vae.single_lane = True
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    for _ in range(2):
        step()
    prof_t = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    wrap.vae.single_lane = False
    ig = prof_t["igemm_tcgen05"]
    peaks = load_peaks()
    achieved = (FLOP_PER_IMAGE if R == 1024 else flops_per_image(R)) * B * 2 / (ig["ms"] * 1e-3) / 1e12 if ig["ms"] else 0.0
    roofline = {
        "bound": "tensor", "kernel": "conv3_fused_kernel + igemm_kernel (tcgen05 implicit GEMM: every conv, projection, QK^T, PV)",
        "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["sustained"],
        "frac_of_burst_peak": achieved / peaks["burst"], "peak_source": peaks["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
        "launches_per_step": ig["launches"] / 2, "avg_launch_ms": ig["ms"] / max(1.0, ig["launches"]),
        "kernel_ms_per_step": ig["ms"] / 2, "algorithmic_flop_per_step": flops_per_image(R) * B,
        "traffic": traffic_per_launch(R, B, ig["launches"] / 2),
        "per_class_ms_per_step": {k: v["ms"] / 2 for k, v in prof_t.items() if v["launches"]},
        "whole_step_frac_of_burst": (value / world) * flops_per_image(R) / 1e12 / peaks["burst"],
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": make_config(B, R, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "vt_infer_host (C-ABI, pinned host buffers)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sample, _ = cpu_reference_rate(30.0, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def traffic_per_launch(R, B, launches_per_step):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per tensor-kernel launch, from the committed
    ``ncu --set full`` capture of one 1024^2 image (profiles/r01_dram_traffic_v20.json) scaled to this step's
    images per launch; None for other resolutions or when the capture is absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_dram_traffic_v20.json")
    if R != 1024 or not os.path.exists(path) or not launches_per_step:
        return None
    with open(path, "r", encoding="utf-8") as f:
        t = json.load(f)["per_image_all_tensor_kernels"]
    return (t["dram_read_MB"] + t["dram_write_MB"]) * 1e6 * B / launches_per_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--resolution", type=int, default=RES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything native code prints there while the benchmark runs (NCCL's
    # version banner under NCCL_DEBUG=VERSION is a plain printf) is sent to stderr at the file-descriptor level;
    # print() inside run_* goes through sys.stdout, which is re-pointed at the real stdout
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
