"""Drop-in for the reference's ``train_full.py``: fine-tune the FLUX VAE encoder AND the tag decoder together
(reference step: train_full.py:201-256).

What runs where:
  * the three VAE encoder forwards of a step (anchor, positive, negative) are native training forwards that keep their
    activations on tape slots, and their backward -- every conv as tcgen05 data- / weight-gradient GEMMs, GroupNorm+SiLU
    and the mid-block attention backward -- is native too (``vt_encoder_train_forward`` / ``vt_encoder_backward``
    behind ``AutoencoderKL.encode`` in ``train()`` mode);
  * the semantic losses on the posterior samples (``ImprovedTripletLoss`` / ``ContrastiveLoss``), the focal loss,
    the reconstruction MSE and the adaptive weighting are native value + gradient kernels (``improved_losses.py``);
  * the tag decoder sees the no-grad posterior mode (train_full.py:217-224) and trains through its PyTorch graph
    here (the fused native head step serves ``train_decoder.py``, where the head is the only trainable part);
  * ``accelerate`` is replaced by plain ``torch.distributed``: one process per GPU, the gradients of both models are
    flattened and all-reduced once per optimizer step.

With ``--use_simplified_loss`` (the reference's default and recommendation) the loss never touches the reconstruction:
it is not computed and only the encoder and the head train.  With the full ``CombinedLoss`` (``--use_full_loss``) the
reconstruction comes from the native decoder TRAINING forward and its MSE is back-propagated natively through the
decoder (parameter gradients) and on into the anchor's posterior sample, i.e. into the encoder.

Launch:  [torchrun --nproc-per-node N] python -m vae_tagger_b200.train_full --json_path ... (the reference's flags).
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
from .improved_losses import CombinedLoss, SimplifiedCombinedLoss, compute_class_distribution
from .modules import ClassificationDecoder, TaggedImageDataset, create_attention_decoder, get_image_transform, \
    get_vae_latent_info
from .train_decoder import _ddp_env, get_scheduler


def _allreduce_grads(params, world):
    """One flat all-reduce (SUM, then / world) over every gradient of the step."""
    grads = [p.grad for p in params if p.grad is not None]
    if world <= 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= world
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def encode_triplet(vae_model, anchor, positive, negative, want_reconstruction):
    """The three VAE passes of train_full.py:210-214.  Returns (reconstruction or None, posteriors, samples)."""
    posts = [vae_model.vae.encode(x).latent_dist for x in (anchor, positive, negative)]
    zs = [p.sample() for p in posts]
    recon = None
    if want_reconstruction:
        recon = vae_model.vae.decode(zs[0]).sample
    return recon, posts, zs


def train_full(args):
    world, rank, local_rank = _ddp_env()
    if not torch.cuda.is_available():
        raise RuntimeError("vae_tagger_b200 needs a CUDA device (B200); there is no CPU path")
    torch.cuda.set_device(local_rank % torch.cuda.device_count())
    device = torch.device("cuda", local_rank % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl")
    random.seed(args.seed + rank)
    torch.manual_seed(args.seed)     # every rank builds the SAME initial models (DDP's start-up broadcast)
    main_proc = rank == 0
    os.makedirs(args.output_dir, exist_ok=True)

    # ---- models (train_full.py:60-118)
    if args.vae_config_path:
        vae_model = create_vae_from_config_file(args.vae_config_path, args.vae_checkpoint)
    else:
        cfg = get_diffusers_vae_config()
        cfg["use_quant_conv"], cfg["use_post_quant_conv"] = args.use_quant_conv, args.use_post_quant_conv
        vae_model = DiffusersVAEWrapper(load_diffusers_vae_from_config(cfg, args.vae_checkpoint))
    vae_model = vae_model.to(device)
    vae_model.vae.precision = "fp32" if args.mixed_precision == "no" else "bf16"
    simplified = args.use_simplified_loss and not args.use_full_loss
    if not simplified:
        vae_model.vae.enable_decoder()    # the reconstruction term trains the VAE decoder too
    for p in vae_model.vae.parameters():
        p.requires_grad_(True)

    tf = get_image_transform(args.resolution)
    dataset = TaggedImageDataset(args.json_path, args.tags_csv_path, tf, use_bucketing=args.use_bucketing,
                                 base_resolution=args.base_resolution, max_resolution=args.max_resolution,
                                 bucket_step=args.bucket_step, triplets=True)
    num_classes = len(dataset.tags)
    info = get_vae_latent_info(args.resolution)
    if args.no_attention:
        decoder = ClassificationDecoder(info["latent_channels"], info["latent_height"], info["latent_width"], num_classes)
    else:
        decoder = create_attention_decoder(
            info["latent_channels"], info["latent_height"], info["latent_width"], num_classes,
            attention_config={"use_spatial_attention": args.use_spatial_attention, "use_self_attention": args.use_self_attention,
                              "use_cross_attention": args.use_cross_attention, "attention_heads": args.attention_heads,
                              "attention_dropout": args.attention_dropout})
    if args.decoder_checkpoint and os.path.exists(args.decoder_checkpoint):
        decoder.load_state_dict(torch.load(args.decoder_checkpoint, map_location="cpu"))
    decoder = decoder.to(device)
    if world > 1:   # identical start on every rank, whatever the checkpoints did
        for t in list(vae_model.vae.parameters()) + list(decoder.parameters()) + list(decoder.buffers()):
            dist.broadcast(t.data, src=0)
    torch.manual_seed(args.seed + 1000 + rank)   # from here on: per-rank randomness (posterior samples, dropout)

    n_val = max(1, int(0.1 * len(dataset))) if len(dataset) > 1 else 0
    g = torch.Generator().manual_seed(args.seed)
    train_set, val_set = torch.utils.data.random_split(dataset, [len(dataset) - n_val, n_val], generator=g)
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, world, rank, shuffle=True) if world > 1 else None
    loader_kw = dict(batch_size=args.train_batch_size, num_workers=args.num_workers, pin_memory=True,
                     persistent_workers=args.num_workers > 0,
                     prefetch_factor=args.prefetch_factor if args.num_workers > 0 else None)
    train_loader = torch.utils.data.DataLoader(train_set, shuffle=sampler is None, sampler=sampler, **loader_kw)
    val_loader = torch.utils.data.DataLoader(val_set, shuffle=False, **loader_kw) if n_val else []
    class_distribution = compute_class_distribution(dataset) if args.use_class_balanced else None

    # ---- loss (train_full.py:141-177)
    if simplified:
        loss_fn = SimplifiedCombinedLoss(classification_weight=args.bce_weight, triplet_weight=args.triplet_weight,
                                         use_focal_loss=args.use_focal_loss, use_class_balanced=args.use_class_balanced,
                                         focal_alpha=args.focal_alpha, focal_gamma=args.focal_gamma,
                                         triplet_margin=args.triplet_margin, similarity_type=args.similarity_type)
    else:
        loss_fn = CombinedLoss(reconstruction_weight=args.reconstruction_weight, kl_weight=args.kl_weight,
                               triplet_weight=args.triplet_weight, classification_weight=args.bce_weight,
                               use_focal_loss=args.use_focal_loss, use_class_balanced=args.use_class_balanced,
                               use_adaptive_weights=args.use_adaptive_weights, focal_alpha=args.focal_alpha,
                               focal_gamma=args.focal_gamma, triplet_margin=args.triplet_margin,
                               similarity_type=args.similarity_type).to(device)
    params = list(vae_model.vae.parameters()) + list(decoder.parameters())
    if not simplified and args.use_adaptive_weights:
        params += list(loss_fn.adaptive_weights.parameters())
    optimizer = torch.optim.AdamW(params, lr=args.learning_rate, weight_decay=args.weight_decay)
    steps_per_epoch = max(1, len(train_loader))
    scheduler = get_scheduler(args.lr_scheduler_type, optimizer, args.lr_warmup_steps, args.num_epochs * steps_per_epoch)

    def run_batch(batch):
        labels = batch["labels"].to(device, non_blocking=True)
        anchor, positive, negative = (batch[k].to(device, non_blocking=True) for k in ("anchor", "positive", "negative"))
        recon, posts, zs = encode_triplet(vae_model, anchor, positive, negative, want_reconstruction=not simplified)
        with torch.no_grad():   # the classifier input is the no-grad posterior mode, scaled like DiffusersVAEWrapper.encode
            latent = posts[0].mode()
            cfg = vae_model.vae.config
            if hasattr(cfg, "scaling_factor"):
                latent = latent * cfg.scaling_factor
            if hasattr(cfg, "shift_factor"):
                latent = latent + cfg.shift_factor
        logits = decoder(latent)
        pos_labels = batch.get("positive_labels", batch["labels"]).to(device, non_blocking=True)
        spc = class_distribution if args.use_class_balanced else None
        if simplified:
            return loss_fn(zs[0], zs[1], zs[2], logits, labels, anchor_labels=labels, positive_labels=pos_labels,
                           samples_per_class=spc)
        return loss_fn(recon, anchor, posts[0], posts[1], posts[2], zs[0], zs[1], zs[2], logits, labels,
                       anchor_labels=labels, positive_labels=pos_labels, samples_per_class=spc)

    history = {"train_loss": [], "val_loss": [], "learning_rates": []}
    best_val = float("inf")
    accum = max(1, args.gradient_accumulation_steps)
    for epoch in range(args.num_epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        vae_model.train()
        decoder.train()
        loss_sum, steps = 0.0, 0
        for step, batch in enumerate(train_loader):
            loss_dict = run_batch(batch)
            total = loss_dict["total_loss"] / accum
            total.backward()
            if (step + 1) % accum == 0:
                _allreduce_grads(params, world)
                if args.max_grad_norm > 0:
                    torch.nn.utils.clip_grad_norm_(params, args.max_grad_norm)
                optimizer.step()
                scheduler.step()
                optimizer.zero_grad(set_to_none=True)
            loss_sum += total.item()
            steps += 1
            if main_proc and step % args.logging_steps == 0:
                parts = ", ".join(f"{k}: {v.item():.4f}" for k, v in loss_dict.items()
                                  if torch.is_tensor(v) and v.numel() == 1 and k != "total_loss")
                print(f"Epoch: {epoch}, Step: {step}, Loss: {total.item():.4f}, {parts}, "
                      f"LR: {optimizer.param_groups[0]['lr']:.2e}")
        vae_model.eval()
        decoder.eval()
        val_sum, val_steps = 0.0, 0
        with torch.no_grad():
            for batch in val_loader:
                val_sum += run_batch(batch)["total_loss"].item()
                val_steps += 1
        avg_train, avg_val = loss_sum / max(1, steps), val_sum / max(1, val_steps)
        history["train_loss"].append(avg_train)
        history["val_loss"].append(avg_val)
        history["learning_rates"].append(optimizer.param_groups[0]["lr"])
        if main_proc:
            print(f"Epoch {epoch} completed - Train Loss: {avg_train:.4f}, Val Loss: {avg_val:.4f}")
            save_now = []
            if avg_val < best_val:
                best_val = avg_val
                save_now.append(("best_vae", "best_decoder"))
            if (epoch + 1) % args.save_steps == 0:
                save_now.append(("vae", "decoder"))
            for vdir, ddir in save_now:
                vae_model.vae.save_pretrained(os.path.join(args.output_dir, vdir))
                os.makedirs(os.path.join(args.output_dir, ddir), exist_ok=True)
                torch.save(decoder.state_dict(), os.path.join(args.output_dir, ddir, "pytorch_model.bin"))
    if main_proc:
        with open(os.path.join(args.output_dir, "training_history.json"), "w") as f:
            json.dump(history, f, indent=2)
    if world > 1 and dist.is_initialized():
        dist.barrier()
    return history


def build_parser():
    p = argparse.ArgumentParser(description="fine-tune the FLUX VAE encoder + tag decoder (B200-native)")
    p.add_argument("--json_path", type=str, required=True)
    p.add_argument("--tags_csv_path", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="full_output")
    p.add_argument("--vae_checkpoint", type=str, default=None)
    p.add_argument("--vae_config_path", type=str, default=None)
    p.add_argument("--decoder_checkpoint", type=str, default=None)
    p.add_argument("--resolution", type=int, default=1024)
    p.add_argument("--train_batch_size", type=int, default=1)
    p.add_argument("--num_epochs", type=int, default=10)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--weight_decay", type=float, default=1e-6)
    p.add_argument("--use_attention", action="store_true", default=True)
    p.add_argument("--no_attention", action="store_true")
    p.add_argument("--use_spatial_attention", action="store_true", default=True)
    p.add_argument("--use_self_attention", action="store_true", default=True)
    p.add_argument("--use_cross_attention", action="store_true")
    p.add_argument("--attention_heads", type=int, default=8)
    p.add_argument("--attention_dropout", type=float, default=0.1)
    p.add_argument("--reconstruction_weight", type=float, default=0.01)
    p.add_argument("--kl_weight", type=float, default=1e-7)
    p.add_argument("--triplet_weight", type=float, default=1.0)
    p.add_argument("--bce_weight", type=float, default=1.0)
    p.add_argument("--triplet_margin", type=float, default=1.0)
    p.add_argument("--use_simplified_loss", action="store_true", default=True)
    p.add_argument("--use_full_loss", action="store_true",
                   help="CombinedLoss instead of the simplified one (the reference's --use_simplified_loss defaults to "
                        "True and cannot be switched off from its command line; this flag can)")
    p.add_argument("--use_focal_loss", action="store_true")
    p.add_argument("--use_class_balanced", action="store_true")
    p.add_argument("--use_adaptive_weights", action="store_true")
    p.add_argument("--focal_alpha", type=float, default=1.0)
    p.add_argument("--focal_gamma", type=float, default=2.0)
    p.add_argument("--similarity_type", type=str, default="cosine", choices=["cosine", "euclidean"])
    p.add_argument("--lr_scheduler_type", type=str, default="cosine")
    p.add_argument("--lr_warmup_steps", type=int, default=500)
    p.add_argument("--max_grad_norm", type=float, default=1.0)
    p.add_argument("--logging_steps", type=int, default=100)
    p.add_argument("--save_steps", type=int, default=5)
    p.add_argument("--mixed_precision", type=str, default="fp16",
                   help="'no' runs the fp32 verification kernels, anything else the 16-bit tensor-core mode")
    p.add_argument("--enable_xformers_memory_efficient_attention", action="store_true", help="accepted and ignored")
    p.add_argument("--use_quant_conv", action="store_true")
    p.add_argument("--use_post_quant_conv", action="store_true")
    p.add_argument("--use_safetensors", action="store_true", help="the VAE is always saved as safetensors")
    p.add_argument("--use_bucketing", action="store_true")
    p.add_argument("--base_resolution", type=int, default=512)
    p.add_argument("--max_resolution", type=int, default=1024)
    p.add_argument("--bucket_step", type=int, default=64)
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--prefetch_factor", type=int, default=2)
    p.add_argument("--gradient_accumulation_steps", type=int, default=1)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--cudnn_benchmark", action="store_true", help="accepted and ignored (no cuDNN on this path)")
    p.add_argument("--cudnn_deterministic", action="store_true", help="accepted and ignored (the kernels are deterministic)")
    return p


def main(argv=None):
    return train_full(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
