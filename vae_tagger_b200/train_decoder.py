"""Drop-in for the reference's ``train_decoder.py``: train the tag decoder on a frozen FLUX VAE
encoder (reference step: train_decoder.py:178-206).

Differences in *how* (not what):
  * the frozen encoder forward -- >99.9 % of the step's FLOPs -- runs as sm_100a kernels;
  * ``accelerate`` is replaced by plain ``torch.distributed`` (one process per GPU, NCCL): decoder
    gradients live in ONE flat fp32 buffer that is all-reduced once per optimizer step; the
    all-reduce is asynchronous and is waited for only after the *next* batch's encoder forward, so
    it is hidden behind it.  BatchNorm statistics stay per rank and the BatchNorm buffers are
    broadcast from rank 0 like DDP's ``broadcast_buffers``;
  * the decoder's train-mode forward, the focal / BCE loss, the backward, gradient clipping and AdamW run
    as native kernels on flat parameter / gradient buffers (``vt_head_train_step``, ``vt_adamw_step``);
    the module's parameters are views into the flat buffer, so checkpoints and ``state_dict`` are
    unchanged (``--use_class_balanced``: the class weights go into the loss kernel; ``--use_cross_attention``: the
    query_generator / CrossAttention branch has its own forward and backward kernels).  A non-AdamW optimizer or
    a custom loss keeps the PyTorch autograd graph for the head only.

Launch:  torchrun --nproc-per-node N -m vae_tagger_b200.train_decoder --vae_checkpoint ... (same flags as
the reference; ``--mixed_precision`` is accepted and ignored: the encoder runs bf16 tensor-core
kernels with fp32 accumulation, the head in fp32).
"""
from __future__ import annotations

import argparse
import json
import math
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _native
from .diffusers_vae_loader import (DiffusersVAEWrapper, create_vae_from_config_file, get_diffusers_vae_config,
                                   load_diffusers_vae_from_config)
from .modules import (AttentionClassificationDecoder, ClassificationDecoder, TaggedImageDataset, create_attention_decoder, get_image_transform,
                      get_vae_latent_info)


def cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps):
    """``get_scheduler("cosine")`` of diffusers: linear warm-up then half a cosine to zero."""

    def lr_lambda(step):
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        progress = (step - num_warmup_steps) / max(1, num_training_steps - num_warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, progress))))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda)


def get_scheduler(name, optimizer, num_warmup_steps, num_training_steps):
    if name == "cosine":
        return cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps)
    if name == "constant":
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: 1.0)
    if name == "linear":
        return torch.optim.lr_scheduler.LambdaLR(
            optimizer, lambda s: s / max(1, num_warmup_steps) if s < num_warmup_steps else max(
                0.0, (num_training_steps - s) / max(1, num_training_steps - num_warmup_steps)))
    raise ValueError(f"unknown lr_scheduler_type {name!r}")


class DecoderTrainer:
    """One training step of the decoder head on a frozen encoder, data-parallel over ``world`` ranks.

    ``native_step``: None = use the native kernels when the configuration has them, True = require them
    (``NativeError`` otherwise), False = PyTorch autograd for the head."""

    def __init__(self, vae_model, decoder, loss_fn, optimizer, scheduler=None, max_grad_norm=1.0,
                 gradient_accumulation_steps=1, process_group=None, native_step=None):
        self.vae, self.decoder, self.loss_fn = vae_model, decoder, loss_fn
        self.opt, self.sched = optimizer, scheduler
        self.max_grad_norm = max_grad_norm
        self.accum = max(1, gradient_accumulation_steps)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in decoder.parameters() if p.requires_grad]
        # one flat gradient bucket; every p.grad is a view into it
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buffers = [b for b in decoder.buffers() if b.is_floating_point()]
        self._buf_pending = []          # in-flight broadcasts of rank 0's BatchNorm buffers
        self._events = None             # enable_timing(): CUDA-event pairs around the collective waits
        self._pending = None
        self._averaged = False
        self._micro = 0
        if self.world > 1:  # initial parameter broadcast (DDP does the same at wrap time)
            for t in list(decoder.parameters()) + list(decoder.buffers()):
                dist.broadcast(t.data, src=0, group=self.pg)
        self.native = False
        why = self._native_plan()
        if native_step is True and why is not None:
            raise _native.NativeError(f"native head training step not available: {why}")
        if native_step is not False and why is None:
            self._go_native()

    # -- native training step -----------------------------------------------------------------
    def _native_plan(self):
        """None when the step can run as native kernels, else the reason it cannot."""
        from .improved_losses import ClassBalancedCriterion, FocalLoss

        dec, dev = self.decoder, self.params[0].device
        if dev.type != "cuda":
            return "decoder is not on a CUDA device"
        if len(self.params) != len(list(dec.parameters())):
            return "some decoder parameters are frozen"
        if isinstance(dec, AttentionClassificationDecoder):
            want_p = (0.3, 0.2, 0.1)
            attn_p = dec.self_attention_post.dropout.p if dec.use_self_attention else 0.0
        elif isinstance(dec, ClassificationDecoder):
            want_p, attn_p = (0.3, 0.2), 0.0
        else:
            return f"{type(dec).__name__} has no kernel"
        ps = tuple(m.p for m in dec.classifier if isinstance(m, nn.Dropout))
        if all(p == 0 for p in ps) and attn_p == 0:
            self._dropout = False
        elif ps == want_p:
            self._dropout = True
        else:
            return f"classifier dropout rates {ps} differ from the reference's {want_p}"
        self._attn_p = float(attn_p)
        self._class_w = None
        if isinstance(self.loss_fn, ClassBalancedCriterion):
            self._alpha, self._gamma = 1.0, 0.0
            self._class_w = torch.tensor(self.loss_fn.weights(), dtype=torch.float32, device=dev)
        elif isinstance(self.loss_fn, FocalLoss) and self.loss_fn.reduction == "mean":
            self._alpha, self._gamma = float(self.loss_fn.alpha), float(self.loss_fn.gamma)
        elif (isinstance(self.loss_fn, nn.BCEWithLogitsLoss) and self.loss_fn.reduction == "mean"
              and self.loss_fn.weight is None and self.loss_fn.pos_weight is None):
            self._alpha, self._gamma = 1.0, 0.0
        else:
            return "loss is none of FocalLoss(mean), BCEWithLogitsLoss(mean), ClassBalancedCriterion"
        o = self.opt
        if type(o) is not torch.optim.AdamW or len(o.param_groups) != 1:
            return "optimizer is not a single-group torch.optim.AdamW"
        g = o.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or len(g["params"]) != len(self.params) or o.state:
            return "AdamW options (amsgrad / maximize / partial parameter list / existing state) have no kernel"
        return None

    def _go_native(self):
        dec, dev = self.decoder, self.params[0].device
        self.ctx = _native.get_context(dev)
        self._head_cfg = dec._head_config()
        self.ctx.configure_head(self._head_cfg[0], **self._head_cfg[1])
        self.ctx._head_owner = None
        layout = self.ctx.head_param_layout()
        named = list(dec.named_parameters())
        if [(n, p.numel()) for n, p in named] != [(n, k) for n, _, k in layout]:
            raise _native.NativeError("decoder.parameters() order differs from vt_head_param_layout")
        total = layout[-1][1] + layout[-1][2]
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        for (_, p), (_, off, n) in zip(named, layout):     # parameters become views of the flat buffer
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self._t = 0
        bn = dec.feature_compress[1] if isinstance(dec, AttentionClassificationDecoder) else None
        self._bn = bn
        self._seed_base = (torch.initial_seed() * 1000003) & (2 ** 63 - 1)
        self.rank = dist.get_rank(self.pg) if self.world > 1 else 0
        self.native = True

    def _native_forward_backward(self, latent, labels):
        self.ctx.configure_head(self._head_cfg[0], **self._head_cfg[1])
        self.ctx._head_owner = None                          # the inference mirror must be re-uploaded
        self.decoder._native_key = None
        bn = self._bn
        loss, _ = self.ctx.head_train_step(
            latent, labels.to(torch.float32), self.flat_param, self.flat_grad,
            bn.running_mean if bn is not None else None, bn.running_var if bn is not None else None,
            bn.num_batches_tracked if bn is not None else None,
            bn_momentum=bn.momentum if bn is not None else 0.1, focal_alpha=self._alpha, focal_gamma=self._gamma,
            loss_scale=1.0 / self.accum, dropout=self._dropout, attention_dropout=self._attn_p,
            seed=self._seed_base + self._micro * self.world + self.rank, class_weights=self._class_w)
        return loss[0]

    # -- exposed time of the collectives ------------------------------------------------------
    def enable_timing(self, on=True):
        """Record CUDA-event pairs on the compute stream around every wait for a collective: the interval is the
        time the compute stream was held up by it (0 when the exchange finished under the encoder forward)."""
        self._events = {"allreduce_wait": [], "buffer_broadcast_wait": []} if on else None

    def _timed_wait(self, key, waits):
        if self._events is None:
            for w in waits:
                w.wait()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for w in waits:
            w.wait()
        e1.record()
        self._events[key].append((e0, e1))

    def timing_summary(self, reset=True):
        """Mean exposed milliseconds per wait, per kind (device time; synchronises)."""
        if self._events is None:
            return {}
        torch.cuda.synchronize()
        out = {k: (sum(a.elapsed_time(b) for a, b in v) / len(v) if v else 0.0) for k, v in self._events.items()}
        if reset:
            self.enable_timing(True)
        return out

    # -- gradient exchange ------------------------------------------------------------------
    def _launch_buffer_broadcast(self):
        """DDP ``broadcast_buffers`` semantics (every forward starts from rank 0's BatchNorm buffers,
        train_decoder.py:194 under accelerate's DDP wrap), issued asynchronously right behind the step that
        produced them: they travel under the next batch's encoder forward instead of in front of the head."""
        if self.world > 1:
            self._buf_pending = [dist.broadcast(b, src=0, group=self.pg, async_op=True) for b in self.buffers]

    def _launch_allreduce(self):
        if self.world > 1:
            self._pending = dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        else:
            self._pending = True

    def finish_update(self):
        """Wait for the in-flight all-reduce (if any), average, clip, step the optimizer."""
        if self._pending is None:
            return
        averaged, self._averaged = self._averaged, False
        if self.world > 1 and not averaged:
            self._timed_wait("allreduce_wait", [self._pending])
            if not self.native:
                self.flat_grad.div_(self.world)
        self._pending = None
        if self.native:    # clip + AdamW + zero_grad in one pass over the flat buffers (averaging folded in)
            g = self.opt.param_groups[0]
            self._t += 1
            self.ctx.adamw_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, lr=g["lr"],
                                betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], step=self._t,
                                grad_scale=1.0 if averaged else 1.0 / self.world, max_norm=float(self.max_grad_norm or 0.0), zero_grad=True)
            self.decoder._native_key = None
            self.opt._opt_called = True                      # the scheduler only checks that a step happened
            if self.sched is not None:
                self.sched.step()
            return
        if self.max_grad_norm and self.max_grad_norm > 0:
            nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
        self.opt.step()
        if self.sched is not None:
            self.sched.step()
        self.flat_grad.zero_()

    # -- one micro-step ----------------------------------------------------------------------
    def step(self, pixel_values, labels):
        with torch.no_grad():
            latent = self.vae.encode(pixel_values)       # frozen encoder; overlaps the pending all-reduce
        self.finish_update()                             # parameters of step k-1 are now final
        if self._buf_pending:                            # rank 0's buffers of step k-1 (sent behind that step)
            self._timed_wait("buffer_broadcast_wait", self._buf_pending)
            self._buf_pending = []
        self.decoder.train()
        # gradient_accumulation_steps > 1, reference semantics (train_decoder.py:193-203): DDP averages every
        # backward, and the ACCUMULATED gradient is clipped after every micro-step, not only before the update
        exact_accum = self.accum > 1
        if exact_accum:
            prev = self.flat_grad.clone()
            self.flat_grad.zero_()
        if self.native:
            loss = self._native_forward_backward(latent, labels)
        else:
            logits = self.decoder(latent)
            loss = self.loss_fn(logits, labels) / self.accum
            loss.backward()                              # accumulates into the flat bucket views
        self._micro += 1
        boundary = self._micro % self.accum == 0
        if exact_accum:
            if self.world > 1:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
                self.flat_grad.div_(self.world)
            self.flat_grad.add_(prev)
            if boundary:
                self._pending, self._averaged = True, True   # clipped (once more) and applied by finish_update
            elif self.max_grad_norm and self.max_grad_norm > 0:
                norm = torch.linalg.vector_norm(self.flat_grad)
                self.flat_grad.mul_(torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0))
        elif boundary:
            self._launch_allreduce()
        self._launch_buffer_broadcast()
        return loss.detach()

    def flush(self):
        self.finish_update()
        for w in self._buf_pending:
            w.wait()
        self._buf_pending = []


def _ddp_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def train_decoder(args):
    import pandas as pd
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    from .improved_losses import ClassBalancedCriterion, FocalLoss, compute_class_distribution

    if not torch.cuda.is_available():
        raise RuntimeError("vae_tagger_b200 needs CUDA devices (B200)")
    world, rank, local = _ddp_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    os.makedirs(args.output_dir, exist_ok=True)
    if args.seed is not None:
        torch.manual_seed(args.seed)

    if args.vae_config_path and os.path.exists(args.vae_config_path):
        vae_model = create_vae_from_config_file(args.vae_config_path, args.vae_checkpoint)
    elif args.vae_checkpoint and os.path.exists(args.vae_checkpoint):
        vae_model = DiffusersVAEWrapper(load_diffusers_vae_from_config(get_diffusers_vae_config(), args.vae_checkpoint))
    else:
        raise RuntimeError("a VAE checkpoint or a VAE config file must be provided")
    vae_model.to(device).eval()
    for p in vae_model.parameters():
        p.requires_grad = False

    info = get_vae_latent_info(args.resolution)
    tags_df = pd.read_csv(args.tags_csv_path)
    num_classes = len(tags_df["name"])
    if args.use_attention:
        decoder = create_attention_decoder(
            info["latent_channels"], info["latent_height"], info["latent_width"], num_classes,
            attention_config={"use_spatial_attention": args.use_spatial_attention,
                              "use_self_attention": args.use_self_attention,
                              "use_cross_attention": args.use_cross_attention,
                              "attention_heads": args.attention_heads,
                              "attention_dropout": args.attention_dropout})
    else:
        decoder = ClassificationDecoder(info["latent_channels"], info["latent_height"], info["latent_width"],
                                        num_classes, use_adaptive_pooling=True)
    if args.decoder_checkpoint and os.path.exists(args.decoder_checkpoint):
        try:
            decoder.load_state_dict(torch.load(args.decoder_checkpoint, map_location="cpu"), strict=False)
        except Exception as e:  # noqa: BLE001 - the reference falls back to training from scratch (:87-92)
            print(f"decoder checkpoint could not be loaded, training from scratch: {e}")
    decoder.to(device)

    gpu_pre = getattr(args, "gpu_preprocess", False)
    transform = None if (args.use_bucketing or gpu_pre) else get_image_transform(args.resolution)
    dataset = TaggedImageDataset(args.json_path, args.tags_csv_path, transform=transform,
                                 use_bucketing=args.use_bucketing, base_resolution=args.base_resolution,
                                 max_resolution=args.max_resolution, bucket_step=args.bucket_step, raw_uint8=gpu_pre)

    def collate_raw(items):  # decoded images differ in size: keep them as a list, stack the labels
        return {"pixel_values": [it["pixel_values"] for it in items], "target_size": [it["target_size"] for it in items],
                "labels": torch.stack([it["labels"] for it in items])}

    def to_device_batch(batch):
        """Device pixel batch of a loader batch: float [B,3,H,W] from the host transform, or -- with --gpu_preprocess --
        uint8 [B,H,W,3] resized on the GPU (SmartResize into the bucket, else the square BILINEAR Resize)."""
        if not gpu_pre:
            return batch["pixel_values"].to(device, non_blocking=True)
        from .preprocess import gpu_smart_resize, gpu_square_resize

        sizes = {t if t is not None else (args.resolution, args.resolution) for t in batch["target_size"]}
        if len(sizes) != 1:
            raise ValueError("a batch mixes aspect-ratio buckets: use --train_batch_size 1 with --use_bucketing "
                             "(images of different buckets cannot be stacked, as in the reference)")
        (w, h), = sizes
        out = torch.empty(len(batch["pixel_values"]), h, w, 3, dtype=torch.uint8, device=device)
        for i, (img, t) in enumerate(zip(batch["pixel_values"], batch["target_size"])):
            img = img.to(device, non_blocking=True)
            if t is not None:
                gpu_smart_resize(img, w, h, out=out[i])
            else:
                gpu_square_resize(img, args.resolution, out=out[i])
        return out

    class_distribution = compute_class_distribution(dataset)
    val_size = max(1, int(len(dataset) * 0.1))
    train_ds, val_ds = torch.utils.data.random_split(dataset, [len(dataset) - val_size, val_size],
                                                     generator=torch.Generator().manual_seed(args.seed or 0))
    sampler = DistributedSampler(train_ds, world, rank, shuffle=True) if world > 1 else None
    collate = collate_raw if gpu_pre else None
    train_dl = DataLoader(train_ds, batch_size=args.train_batch_size, shuffle=sampler is None, sampler=sampler,
                          pin_memory=True, num_workers=args.num_workers, collate_fn=collate,
                          prefetch_factor=args.prefetch_factor if args.num_workers > 0 else None,
                          persistent_workers=args.num_workers > 0)
    val_dl = DataLoader(val_ds, batch_size=args.train_batch_size, shuffle=False, pin_memory=True, num_workers=0,
                        collate_fn=collate)

    loss_fn = FocalLoss(alpha=args.focal_alpha, gamma=args.focal_gamma) if args.use_focal_loss else nn.BCEWithLogitsLoss()
    # train_decoder.py:188-189: with --use_class_balanced the class-balanced loss replaces the focal / BCE one
    base_fn = ClassBalancedCriterion(class_distribution) if args.use_class_balanced else loss_fn
    optimizer = torch.optim.AdamW(decoder.parameters(), lr=args.learning_rate, weight_decay=args.weight_decay)
    scheduler = get_scheduler(args.lr_scheduler_type, optimizer, args.lr_warmup_steps, args.num_epochs * len(train_dl))
    trainer = DecoderTrainer(vae_model, decoder, base_fn, optimizer, scheduler, args.max_grad_norm,
                             args.gradient_accumulation_steps)

    best, history = float("inf"), {"train_loss": [], "val_loss": [], "learning_rates": []}
    for epoch in range(args.num_epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        total, steps = torch.zeros((), device=device), 0
        for step, batch in enumerate(train_dl):
            loss = trainer.step(to_device_batch(batch), batch["labels"].to(device, non_blocking=True))
            total += loss
            steps += 1
            if rank == 0 and step % args.logging_steps == 0:  # the only host sync of the loop
                print(f"Epoch: {epoch}, Step: {step}, Loss: {loss.item():.4f}, "
                      f"Avg Loss: {(total / steps).item():.4f}, LR: {optimizer.param_groups[0]['lr']:.2e}")
        trainer.flush()
        decoder.eval()
        vtotal, vsteps = torch.zeros((), device=device), 0
        with torch.no_grad():
            for batch in val_dl:
                logits = decoder(vae_model.encode(to_device_batch(batch)))
                vtotal += base_fn(logits, batch["labels"].to(device))
                vsteps += 1
        tl, vl = (total / max(1, steps)).item(), (vtotal / max(1, vsteps)).item()
        history["train_loss"].append(tl); history["val_loss"].append(vl)
        history["learning_rates"].append(optimizer.param_groups[0]["lr"])
        if rank == 0:
            print(f"Epoch {epoch} completed - Train Loss: {tl:.4f}, Val Loss: {vl:.4f}")
            if vl < best:
                best = vl
                torch.save(decoder.state_dict(), os.path.join(args.output_dir, "best_pytorch_model.bin"))
            if (epoch + 1) % args.save_steps == 0:
                torch.save(decoder.state_dict(), os.path.join(args.output_dir, "pytorch_model.bin"))
    if rank == 0:
        with open(os.path.join(args.output_dir, "training_history.json"), "w") as f:
            json.dump(history, f, indent=2)
    if world > 1:
        dist.barrier()
    return history


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--vae_checkpoint", type=str, required=True)
    p.add_argument("--vae_config_path", type=str, default=None)
    p.add_argument("--decoder_checkpoint", type=str, default=None)
    p.add_argument("--json_path", type=str, required=True)
    p.add_argument("--tags_csv_path", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="decoder_output")
    p.add_argument("--resolution", type=int, default=1024)
    p.add_argument("--train_batch_size", type=int, default=1)
    p.add_argument("--num_epochs", type=int, default=10)
    p.add_argument("--learning_rate", type=float, default=1e-3)
    p.add_argument("--weight_decay", type=float, default=1e-6)
    p.add_argument("--mixed_precision", type=str, default="fp16")
    p.add_argument("--use_attention", action="store_true", default=True)
    p.add_argument("--no_attention", action="store_true")
    p.add_argument("--use_spatial_attention", action="store_true", default=True)
    p.add_argument("--use_self_attention", action="store_true", default=True)
    p.add_argument("--use_cross_attention", action="store_true")
    p.add_argument("--attention_heads", type=int, default=8)
    p.add_argument("--attention_dropout", type=float, default=0.1)
    p.add_argument("--use_simplified_decoder_loss", action="store_true", default=True)
    p.add_argument("--use_focal_loss", action="store_true")
    p.add_argument("--use_class_balanced", action="store_true")
    p.add_argument("--focal_alpha", type=float, default=1.0)
    p.add_argument("--focal_gamma", type=float, default=2.0)
    p.add_argument("--lr_scheduler_type", type=str, default="cosine")
    p.add_argument("--lr_warmup_steps", type=int, default=500)
    p.add_argument("--max_grad_norm", type=float, default=1.0)
    p.add_argument("--logging_steps", type=int, default=100)
    p.add_argument("--save_steps", type=int, default=5)
    p.add_argument("--use_quant_conv", action="store_true")
    p.add_argument("--use_post_quant_conv", action="store_true")
    p.add_argument("--use_safetensors", action="store_true")
    p.add_argument("--use_bucketing", action="store_true")
    p.add_argument("--base_resolution", type=int, default=512)
    p.add_argument("--max_resolution", type=int, default=1024)
    p.add_argument("--bucket_step", type=int, default=64)
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--prefetch_factor", type=int, default=2)
    p.add_argument("--gradient_accumulation_steps", type=int, default=1)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--cudnn_benchmark", action="store_true")
    p.add_argument("--cudnn_deterministic", action="store_true")
    # addition of this implementation (default keeps the reference's host-side PIL transform)
    p.add_argument("--gpu_preprocess", action="store_true",
                   help="loader workers only decode; resize / crop / normalise run on the GPU (bit-exact with PIL)")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.no_attention:
        args.use_attention = False
    return train_decoder(args)


if __name__ == "__main__":
    main()
