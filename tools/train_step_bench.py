"""BASELINE config 5: train_decoder step with a frozen encoder (1024^2, bf16 encoder kernels) -- frozen
encoder forward + head train-mode forward/backward (focal loss) + gradient all-reduce (N > 1) + clip +
AdamW.  Times the whole step and, separately, the head-only part with the native kernels
(vt_head_train_step + vt_adamw_step) and with the PyTorch autograd graph of the same module.

    python tools/train_step_bench.py [--batch 8] [--res 1024] [--tags 1000] [--steps 5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/train_step_bench.py
Prints one JSON line (rank 0).  Device time: CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402
from vae_tagger_b200.improved_losses import FocalLoss  # noqa: E402
from vae_tagger_b200.train_decoder import DecoderTrainer  # noqa: E402


def timed(fn, steps, warmup, world):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--tags", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).to(dev).eval()
    for p in wrap.parameters():
        p.requires_grad = False
    g = torch.Generator().manual_seed(7 + rank)
    x = (torch.rand(a.batch, 3, a.res, a.res, generator=g) * 2 - 1).to(dev)
    y = (torch.rand(a.batch, a.tags, generator=g) < 0.1).float().to(dev)
    out = {"metric": "train_decoder step (frozen encoder fwd + head fwd/bwd + all-reduce + AdamW)", "n_gpus": world,
           "batch_per_gpu": a.batch, "resolution": a.res, "tags": a.tags, "steps": a.steps}

    class Frozen(torch.nn.Module):
        def __init__(self, lat):
            super().__init__()
            self.lat = lat

        def encode(self, _):
            return self.lat

    with torch.no_grad():
        lat = wrap.encode(x)
    for native in (True, False):
        for name, vae in (("step", wrap), ("head_only", Frozen(lat))):
            torch.manual_seed(1)
            dec = M.create_attention_decoder(16, a.res // 8, a.res // 8, a.tags, attention_config={}).to(dev)
            opt = torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-6)
            tr = DecoderTrainer(vae, dec, FocalLoss(1.0, 2.0), opt, None, max_grad_norm=1.0, native_step=native)
            ms = timed(lambda: tr.step(x, y), a.steps, a.warmup, world)
            tr.flush()
            out[f"{name}_ms_{'native' if native else 'autograd'}"] = round(ms, 3)
    out["images_per_s_native"] = round(world * a.batch / out["step_ms_native"] * 1e3, 2)
    out["images_per_s_autograd_head"] = round(world * a.batch / out["step_ms_autograd"] * 1e3, 2)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
