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


def kernel_source_sha():
    """sha256 over the kernel / C-ABI sources (vae_tagger_b200/csrc/*, include/*.h, sorted by name): identifies the
    BUILD a profile was taken on, independent of commits that do not touch a kernel."""
    import hashlib

    h = hashlib.sha256()
    for d in (os.path.join(ROOT, "vae_tagger_b200", "csrc"), os.path.join(ROOT, "include")):
        for name in sorted(os.listdir(d)):
            if name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                with open(os.path.join(d, name), "rb") as f:
                    h.update(f.read())
    return h.hexdigest()


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
    cfg = make_config(args.batch, args.resolution, max(1, args.gpus))
    # what this arm actually times: ONE image per step on the host cores of rank 0 (a bounded sample of the
    # workload above, scaled to 1024^2 when a smaller image had to be used); no GPU, one process at any N
    cfg["global_batch"] = 1
    cfg["parallelism"] = "1 CPU process (rank 0), all host threads"
    cfg["workload"] = ("bounded sample of: " + cfg["workload"] + " -- timed here: batch 1 (one image per step) through "
                       "the CPU restatement of the reference path (oracle/), fp32")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": warm, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
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
    peaks = load_peaks()
    tensor_cls = [k for k in _native.TENSOR_KERNEL_CLASSES if prof_t[k]["launches"]]
    t_ms = sum(prof_t[k]["ms"] for k in tensor_cls)
    t_launches = sum(prof_t[k]["launches"] for k in tensor_cls)
    alg_flop = FLOP_PER_IMAGE if R == 1024 else flops_per_image(R)
    achieved = alg_flop * B * 2 / (t_ms * 1e-3) / 1e12 if t_ms else 0.0
    # every tensor kernel family against its own roofline (its own algorithmic FLOPs / bytes as the launchers
    # count them, CUDA events around each launch, inside a full step: sustained peak)
    per_kernel = {}
    for k in tensor_cls:
        v = prof_t[k]
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        hbm_bound = k == "conv_in"      # K = 27: 12 B read + 256 B written per pixel, 7 kFLOP -- write bound
        per_kernel[k] = {"launches_per_step": v["launches"] / 2, "ms_per_step": v["ms"] / 2, "share_of_tensor_ms": v["ms"] / t_ms,
                         "bound": "hbm" if hbm_bound else "tensor",
                         "achieved": gbs if hbm_bound else tf, "unit": "GB/s" if hbm_bound else "TFLOP/s",
                         "frac": (gbs / peaks["hbm"]) if hbm_bound else (tf / peaks["sustained"])}
    roofline = {
        "bound": "tensor", "kernel": "all tcgen05 contraction kernels of the step (conv3_fused_kernel x3 variants, igemm_kernel, "
                                     "flash_d512_kernel, conv_in_kernel); per family in per_kernel",
        "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["sustained"],
        "frac_of_burst_peak": achieved / peaks["burst"], "peak_source": peaks["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
        "launches_per_step": t_launches / 2, "avg_launch_ms": t_ms / max(1.0, t_launches),
        "kernel_ms_per_step": t_ms / 2, "algorithmic_flop_per_step": flops_per_image(R) * B,
        "per_kernel": per_kernel,
        "per_class_ms_per_step": {k: v["ms"] / 2 for k, v in prof_t.items() if v["launches"]},
        "whole_step_frac_of_burst": (value / world) * flops_per_image(R) / 1e12 / peaks["burst"],
    }
    roofline.update(traffic_per_launch(R, B, t_launches / 2))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("16-bit tensor-core mode: bf16 storage of raw activations (VT_B200_RAW_BF16=1), fp16 bounded MMA "
                  "operands and weights, fp32 accumulation in TMEM"
                  if os.environ.get("VT_B200_RAW_BF16", "0")[:1] == "1" else
                  "16-bit tensor-core mode ('bf16 mode' of the north star): fp16 storage and MMA operands -- the "
                  "reference's own autocast dtype (infer_full.py:100), same tensor peak as bf16 -- fp32 accumulation "
                  "in TMEM; VT_B200_RAW_BF16=1 stores raw activations as bf16 instead (same speed, 5x the error)"),
        "data": "synthetic",
        "config": make_config(B, R, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "vt_infer_host (C-ABI, pinned host buffers)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    # ---- auxiliary blocks beside the (unchanged) timed headline: configs[4] training step with the exposed
    # all-reduce time, configs[3] bulk stream; default at N > 1 so that the scaling record carries them
    if args.aux == "on" or (args.aux == "auto" and world > 1):
        line["train_step"] = aux_train_step(world, rank, dev, wrap)
        # the training block loaded its own head (and, below, the fine-tuning block its own encoder) into the
        # device's context: hand the context back to the benched modules before the stream runs through it
        wrap.vae._sync_native(dev)
        dec._native_ctx(dev)
        line["bulk_stream"] = aux_bulk_stream(world, rank, dev, ctx, images_per_gpu=args.bulk_images)
    if args.aux != "off":
        line["encoder_train"] = aux_encoder_train(world, rank, dev)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sample, _ = cpu_reference_rate(30.0, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def aux_train_step(world, rank, dev, wrap, steps=6, warmup=3, batch=8, res=RES):
    """BASELINE configs[4] / SURVEY 8d-5 beside the headline (not part of the timed region): train_decoder step with
    the frozen encoder -- encoder fwd (batch per GPU, 1024^2, 16-bit kernels) + head fwd/bwd (focal loss, native
    kernels) + NCCL all-reduce of the flat fp32 gradient + clip + AdamW.  Device time (CUDA events), max over
    ranks; the exposed time of a collective is the interval the compute stream waited for it."""
    import torch
    import torch.distributed as dist

    from vae_tagger_b200 import modules as M
    from vae_tagger_b200.improved_losses import FocalLoss
    from vae_tagger_b200.train_decoder import DecoderTrainer

    g = torch.Generator().manual_seed(7 + rank)
    x = (torch.rand(batch, 3, res, res, generator=g) * 2 - 1).to(dev)
    y = (torch.rand(batch, NUM_TAGS, generator=g) < 0.1).float().to(dev)

    class Frozen(torch.nn.Module):
        def __init__(self, lat):
            super().__init__()
            self.lat = lat

        def encode(self, _):
            return self.lat

    def run(vae):
        torch.manual_seed(1)
        dec = M.create_attention_decoder(16, res // 8, res // 8, NUM_TAGS, attention_config={}).to(dev)
        opt = torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-6)
        tr = DecoderTrainer(vae, dec, FocalLoss(1.0, 2.0), opt, None, max_grad_norm=1.0, native_step=True)
        for _ in range(warmup):
            tr.step(x, y)
        tr.enable_timing(True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.step(x, y)
        e1.record()
        torch.cuda.synchronize()
        exposed = tr.timing_summary()
        tr.flush()
        t = torch.tensor([e0.elapsed_time(e1) / steps, exposed.get("allreduce_wait", 0.0),
                          exposed.get("buffer_broadcast_wait", 0.0)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist(), sum(p.numel() for p in dec.parameters())

    with torch.no_grad():
        lat = wrap.encode(x)
    (step_ms, ar_ms, bc_ms), nparam = run(wrap)
    (head_ms, _, _), _ = run(Frozen(lat))
    return {"workload": f"configs[4]: train_decoder step, frozen encoder, {res}x{res}, batch {batch} per GPU, {NUM_TAGS} tags, "
                        "focal loss, AdamW, native head kernels", "n_gpus": world, "steps": steps,
            "step_ms": step_ms, "head_fwd_bwd_adamw_ms": head_ms, "allreduce_exposed_ms": ar_ms,
            "buffer_broadcast_exposed_ms": bc_ms, "allreduce_bytes": 4 * nparam,
            "images_per_s": world * batch / step_ms * 1e3,
            "how": "CUDA events on the compute stream, max over ranks; exposed = event interval around the wait for the "
                   "async NCCL op (launched behind step k, waited after the encoder forward of step k+1)"}


def aux_encoder_train(world, rank, dev, batch=2, res=512, steps=3, warmup=2):
    """SURVEY 8f-4 beside the headline (not part of the timed region): one VAE fine-tuning pass of the encoder --
    native training forward (activations kept on a tape) + native backward (all 106 parameter gradients), the work
    autograd does for the reference in train_full.py:201-256.  Device time (CUDA events), max over ranks."""
    import torch
    import torch.distributed as dist

    from vae_tagger_b200 import diffusers_vae_loader as L

    torch.manual_seed(3)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config()).to(dev).train()
    x = (torch.rand(batch, 3, res, res, generator=torch.Generator().manual_seed(11 + rank)) * 2 - 1).to(dev)
    nctx = vae._sync_native(dev)
    grads = {n: torch.empty_like(p, dtype=torch.float32) for n, p in vae.encoder.named_parameters()}
    gm = torch.randn(batch, 16, res // 8, res // 8, device=dev)

    def step():
        nctx.encode_train(x, precision=vae._precision(), slot=0)
        nctx.encoder_backward(gm, None, grads, slot=0)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nctx.release_tape(0)
    ms = t.item()
    flop = 3.0 * flops_per_image(res) * batch      # forward + data gradients + weight gradients
    return {"workload": f"encoder fine-tuning pass (training forward + backward), {batch} x {res}x{res} per GPU, 16-bit mode",
            "n_gpus": world, "steps": steps, "step_ms": ms, "images_per_s": world * batch / ms * 1e3,
            "algorithmic_tflops": flop / ms / 1e9, "parameter_gradients": len(grads),
            "how": "CUDA events around vt_encoder_train_forward + vt_encoder_backward, max over ranks"}


def aux_bulk_stream(world, rank, dev, ctx, images_per_gpu=8192, batch=BATCH, res=RES):
    """BASELINE configs[3] / SURVEY 8d-4 beside the headline: batch-sharded bulk tagging of a synthetic uint8
    image stream (3 MB per image, pinned host memory, double buffered) through vt_infer_host; rank r owns a
    contiguous shard, no collective on the data path.  Wall clock, max over ranks."""
    import torch
    import torch.distributed as dist

    g = torch.Generator().manual_seed(1000 + rank)
    pool = [torch.randint(0, 256, (batch, res, res, 3), generator=g, dtype=torch.uint8) for _ in range(2)]
    bufs = [torch.empty(batch, res, res, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
    outs = [None, None]
    nb = max(1, images_per_gpu // batch)
    tags = 0

    def fill(slot, k):   # stands in for decode + collate of the next batch
        bufs[slot].copy_(pool[k & 1])

    fill(0, 0)
    ctx.infer_host(bufs[0], threshold=0.5)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(nb):
        slot = k & 1
        th = None
        if k + 1 < nb:
            th = threading.Thread(target=fill, args=(slot ^ 1, k + 1))
            th.start()
        outs[slot] = ctx.infer_host(bufs[slot], threshold=0.5, out=outs[slot])
        tags += int(outs[slot]["count"].sum())
        if th is not None:
            th.join()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = nb * batch
    return {"workload": f"configs[3]: bulk tagging of a uint8 host stream, {res}x{res}, {n} images per GPU "
                        f"({n * world} total), contiguous shard per rank", "n_gpus": world, "images": n * world,
            "images_per_s": n * world / t.item(), "seconds": t.item(), "h2d_bytes_per_image": res * res * 3,
            "collectives_on_data_path": 0, "mean_tags_above_threshold": tags / n}


def traffic_per_launch(R, B, launches_per_step):
    """``roofline.traffic``: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per tensor-kernel launch, from
    the committed ``ncu --set full`` capture of one 1024^2 image (profiles/r02_dram_traffic.json, written by
    tools/make_traffic_json.py on the GPU box) scaled to this step's images per launch.  The capture carries the
    sha256 of the kernel sources it was taken on; when that differs from the sources of THIS build the number is
    not printed (null) -- a traffic figure of another build is not evidence for this one."""
    path = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    out = {"traffic": None, "traffic_source": None}
    if R != 1024 or not os.path.exists(path) or not launches_per_step:
        return out
    with open(path, "r", encoding="utf-8") as f:
        d = json.load(f)
    sha = kernel_source_sha()
    if d.get("kernel_source_sha256") != sha:
        out["traffic_source"] = (f"profiles/r02_dram_traffic.json was captured on kernel sources {str(d.get('kernel_source_sha256'))[:12]}, "
                                 f"this build is {sha[:12]}: not reported")
        return out
    t = d["per_image_all_tensor_kernels"]
    out["traffic"] = (t["dram_read_MB"] + t["dram_write_MB"]) * 1e6 * B / launches_per_step
    out["traffic_source"] = f"profiles/r02_dram_traffic.json (kernel sources {sha[:12]}, same as this build)"
    out["traffic_per_image_by_kernel_MB"] = {k: round(v["dram_read_MB"] + v["dram_write_MB"], 1)
                                             for k, v in d["per_image_by_kernel"].items()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--resolution", type=int, default=RES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--aux", default="auto", choices=["auto", "on", "off"],
                    help="train_step / bulk_stream blocks in the JSON line (auto: only under torchrun, N > 1)")
    ap.add_argument("--bulk-images", type=int, default=8192, help="images per GPU of the bulk_stream block")
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
