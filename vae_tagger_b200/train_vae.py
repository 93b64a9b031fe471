"""Drop-in for the reference's ``train_vae.py``: fine-tune the whole FLUX VAE (encoder AND decoder) with a
reconstruction + (log-stabilised KL) + triplet objective (reference step: train_vae.py:124-186).

Per step: ``model(anchor)`` = native encoder training forward, posterior sample, native decoder training forward;
``model(positive)`` / ``model(negative)`` likewise (the reference decodes those two as well and throws the images
away -- here only their encoders run: the losses never touch those reconstructions); ``F.mse_loss`` and
``ImprovedTripletLoss`` are native value + gradient kernels; ``.backward()`` runs the native decoder backward
(``d loss / d image`` -> decoder parameter gradients + ``d loss / d z``) and the three native encoder backwards.
``accelerate`` is replaced by plain ``torch.distributed`` (one flat gradient all-reduce per step).

Launch:  [torchrun --nproc-per-node N] python -m vae_tagger_b200.train_vae --json_path ... (the reference's flags).
"""
from __future__ import annotations

import argparse
import json
import os
import random

import torch
import torch.distributed as dist

from .diffusers_vae_loader import DiffusersVAEWrapper, create_vae_from_config_file, get_diffusers_vae_config, \
    load_diffusers_vae_from_config
from .improved_losses import ImprovedTripletLoss, mse_loss
from .modules import TaggedImageDataset, get_image_transform
from .train_decoder import _ddp_env, get_scheduler
from .train_full import _allreduce_grads


def vae_step_losses(model, batch, device, triplet_loss_fn, args):
    """The loss terms of train_vae.py:124-186 for one batch.  Returns (total, recon, kl_for_log, triplet)."""
    anchor, positive, negative = (batch[k].to(device, non_blocking=True) for k in ("anchor", "positive", "negative"))
    anchor_labels = batch["labels"].to(device, non_blocking=True)
    positive_labels = batch.get("positive_labels", batch["labels"]).to(device, non_blocking=True)
    reconstruction_a, posterior_a = model(anchor)
    posterior_p = model.vae.encode(positive).latent_dist
    posterior_n = model.vae.encode(negative).latent_dist
    z_a, z_p, z_n = posterior_a.sample(), posterior_p.sample(), posterior_n.sample()
    recon_loss = mse_loss(reconstruction_a, anchor)
    triplet_loss = triplet_loss_fn(z_a.reshape(z_a.size(0), -1), z_p.reshape(z_p.size(0), -1), z_n.reshape(z_n.size(0), -1),
                                   anchor_labels, positive_labels)
    kl_mean = ((posterior_a.kl() + posterior_p.kl() + posterior_n.kl()) / 3).mean()
    kl_loss = torch.log(1 + kl_mean / 10000)
    if args.use_simplified_vae_loss and not args.use_kl_loss:
        total = args.reconstruction_weight * recon_loss + args.triplet_weight * triplet_loss
    else:
        total = args.reconstruction_weight * recon_loss + args.kl_weight * kl_loss + args.triplet_weight * triplet_loss
    return total, recon_loss, kl_loss, triplet_loss


def train_vae(args):
    world, rank, local_rank = _ddp_env()
    if not torch.cuda.is_available():
        raise RuntimeError("vae_tagger_b200 needs a CUDA device (B200); there is no CPU path")
    torch.cuda.set_device(local_rank % torch.cuda.device_count())
    device = torch.device("cuda", local_rank % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl")
    random.seed(args.seed + rank)
    torch.manual_seed(args.seed)
    main_proc = rank == 0
    os.makedirs(args.output_dir, exist_ok=True)

    if args.vae_config_path and os.path.exists(args.vae_config_path):
        model = create_vae_from_config_file(args.vae_config_path, args.vae_checkpoint)
    else:
        cfg = get_diffusers_vae_config()
        if not (args.vae_checkpoint and os.path.exists(args.vae_checkpoint)):
            cfg["sample_size"] = args.resolution
        model = DiffusersVAEWrapper(load_diffusers_vae_from_config(cfg, args.vae_checkpoint))
    model = model.to(device)
    model.vae.enable_decoder()
    model.vae.precision = "fp32" if args.mixed_precision == "no" else "bf16"
    params = list(model.vae.parameters())
    for p in params:
        p.requires_grad_(True)
    if world > 1:
        for p in params:
            dist.broadcast(p.data, src=0)
    torch.manual_seed(args.seed + 1000 + rank)

    tf = None if args.use_bucketing else get_image_transform(args.resolution)
    dataset = TaggedImageDataset(args.json_path, args.tags_csv_path, tf, use_bucketing=args.use_bucketing,
                                 base_resolution=args.base_resolution, max_resolution=args.max_resolution,
                                 bucket_step=args.bucket_step, triplets=True)
    n_val = max(1, int(0.1 * len(dataset)))
    g = torch.Generator().manual_seed(args.seed)
    train_set, val_set = torch.utils.data.random_split(dataset, [len(dataset) - n_val, n_val], generator=g)
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, world, rank, shuffle=True) if world > 1 else None
    loader_kw = dict(batch_size=args.train_batch_size, num_workers=args.num_workers, pin_memory=True,
                     persistent_workers=args.num_workers > 0,
                     prefetch_factor=args.prefetch_factor if args.num_workers > 0 else None)
    train_loader = torch.utils.data.DataLoader(train_set, shuffle=sampler is None, sampler=sampler, **loader_kw)
    val_loader = torch.utils.data.DataLoader(val_set, shuffle=False, **loader_kw)

    triplet_loss_fn = ImprovedTripletLoss(margin=args.triplet_margin, similarity_type=args.similarity_type)
    optimizer = torch.optim.AdamW(params, lr=args.learning_rate, weight_decay=args.weight_decay)
    scheduler = get_scheduler(args.lr_scheduler_type, optimizer, args.lr_warmup_steps,
                              args.num_epochs * max(1, len(train_loader)))

    history = {"train_loss": [], "val_loss": [], "learning_rates": []}
    best_val = float("inf")
    for epoch in range(args.num_epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        model.train()
        loss_sum, steps = 0.0, 0
        for step, batch in enumerate(train_loader):
            total, recon, kl, trip = vae_step_losses(model, batch, device, triplet_loss_fn, args)
            total.backward()
            _allreduce_grads(params, world)
            if args.max_grad_norm > 0:
                torch.nn.utils.clip_grad_norm_(params, args.max_grad_norm)
            optimizer.step()
            scheduler.step()
            optimizer.zero_grad(set_to_none=True)
            loss_sum += total.item()
            steps += 1
            if main_proc and step % args.logging_steps == 0:
                print(f"Epoch: {epoch}, Step: {step}, Total: {total.item():.4f}, Recon: {recon.item():.4f}, "
                      f"KL: {kl.item():.4f}, Triplet: {trip.item():.4f}, LR: {optimizer.param_groups[0]['lr']:.2e}")
        model.eval()
        val_sum, val_steps = 0.0, 0
        with torch.no_grad():
            for batch in val_loader:
                val_sum += vae_step_losses(model, batch, device, triplet_loss_fn, args)[0].item()
                val_steps += 1
        avg_train, avg_val = loss_sum / max(1, steps), val_sum / max(1, val_steps)
        history["train_loss"].append(avg_train)
        history["val_loss"].append(avg_val)
        history["learning_rates"].append(optimizer.param_groups[0]["lr"])
        if main_proc:
            print(f"Epoch {epoch} completed - Train Loss: {avg_train:.4f}, Val Loss: {avg_val:.4f}")
            if avg_val < best_val:
                best_val = avg_val
                model.vae.save_pretrained(os.path.join(args.output_dir, "best_vae"))
            if (epoch + 1) % args.save_steps == 0:
                model.vae.save_pretrained(os.path.join(args.output_dir, f"vae_checkpoint_epoch_{epoch}"))
    if main_proc:
        with open(os.path.join(args.output_dir, "training_history.json"), "w") as f:
            json.dump(history, f, indent=2)
    if world > 1 and dist.is_initialized():
        dist.barrier()
    return history


def build_parser():
    p = argparse.ArgumentParser(description="fine-tune the FLUX VAE (B200-native)")
    p.add_argument("--vae_checkpoint", type=str, default=None)
    p.add_argument("--vae_config_path", type=str, default=None)
    p.add_argument("--json_path", type=str, required=True)
    p.add_argument("--tags_csv_path", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="vae_output")
    p.add_argument("--resolution", type=int, default=1024)
    p.add_argument("--train_batch_size", type=int, default=1)
    p.add_argument("--num_epochs", type=int, default=10)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--weight_decay", type=float, default=1e-6)
    p.add_argument("--use_simplified_vae_loss", action="store_true", default=True)
    p.add_argument("--use_kl_loss", action="store_true",
                   help="reconstruction + KL + triplet (the reference's --use_simplified_vae_loss defaults to True and "
                        "cannot be switched off from its command line; this flag selects its other branch)")
    p.add_argument("--reconstruction_weight", type=float, default=0.01)
    p.add_argument("--kl_weight", type=float, default=1e-2)
    p.add_argument("--triplet_weight", type=float, default=1.0)
    p.add_argument("--triplet_margin", type=float, default=1.0)
    p.add_argument("--similarity_type", type=str, default="cosine", choices=["cosine", "euclidean"])
    p.add_argument("--lr_scheduler_type", type=str, default="cosine")
    p.add_argument("--lr_warmup_steps", type=int, default=500)
    p.add_argument("--max_grad_norm", type=float, default=1.0)
    p.add_argument("--logging_steps", type=int, default=100)
    p.add_argument("--save_steps", type=int, default=5)
    p.add_argument("--mixed_precision", type=str, default="fp16",
                   help="'no' runs the fp32 verification kernels, anything else the 16-bit tensor-core mode")
    p.add_argument("--enable_xformers_memory_efficient_attention", action="store_true", help="accepted and ignored")
    p.add_argument("--use_bucketing", action="store_true")
    p.add_argument("--base_resolution", type=int, default=512)
    p.add_argument("--max_resolution", type=int, default=1024)
    p.add_argument("--bucket_step", type=int, default=64)
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--prefetch_factor", type=int, default=2)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--cudnn_benchmark", action="store_true", help="accepted and ignored (no cuDNN on this path)")
    p.add_argument("--cudnn_deterministic", action="store_true", help="accepted and ignored (the kernels are deterministic)")
    return p


def main(argv=None):
    return train_vae(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
