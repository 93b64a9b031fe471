"""Drop-in for the reference's ``infer_vae.py``: encode images with the FLUX VAE encoder and dump the
flattened latents (``mode()*scaling_factor + shift_factor``) to ``latent_vectors.json`` -- same CLI flags
and output schema (reference :31-81), batched through the sm_100a encoder with one device->host copy per
batch.

    python -m vae_tagger_b200.infer_vae --vae_checkpoint vae.safetensors --vae_config_path cfg.json --image_path imgs/
"""
from __future__ import annotations

import argparse
import json
import os
from pathlib import Path

import torch

from .diffusers_vae_loader import (DiffusersVAEWrapper, create_vae_from_config_file, get_diffusers_vae_config,
                                   load_diffusers_vae_from_config)
from .modules import get_image_paths, get_image_transform


def load_vae(args, device="cuda"):
    if args.vae_config_path and os.path.exists(args.vae_config_path):
        model = create_vae_from_config_file(args.vae_config_path, args.vae_checkpoint)
    elif args.vae_checkpoint and os.path.exists(args.vae_checkpoint):
        model = DiffusersVAEWrapper(load_diffusers_vae_from_config(get_diffusers_vae_config(), args.vae_checkpoint))
    else:
        raise RuntimeError("a VAE checkpoint or a VAE config file must be provided")
    return model.to(device).eval()


@torch.no_grad()
def infer_and_save_latents(args):
    from PIL import Image

    from .sharding import dist_env, gather_to_rank0, init_host_group, shard_range

    if not torch.cuda.is_available():
        raise RuntimeError("vae_tagger_b200 needs a CUDA device (B200); there is no CPU path")
    # under torchrun: rank r encodes a contiguous shard of the image list on its own GPU (no collective), the
    # latents are gathered on rank 0's host and rank 0 writes ONE latent_vectors.json
    rank, world, local_rank = dist_env()
    if world > 1:
        local_rank %= torch.cuda.device_count()     # more ranks than GPUs (a test box): ranks share a device
        torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}" if world > 1 else "cuda"
    own_group = init_host_group(world)
    vae_model = load_vae(args, device)
    transform = get_image_transform(args.resolution)
    if not os.path.exists(args.image_path):
        raise FileNotFoundError(f"image path not found: {args.image_path}")
    image_paths = get_image_paths(args.image_path)
    if not image_paths:
        print("no image files found")
        return {}
    latent_data, errors = {}, 0
    bs = max(1, getattr(args, "batch_size", 8))
    lo, hi = shard_range(len(image_paths), rank, world)
    for i0 in range(lo, hi, bs):
        tensors, names = [], []
        for p in image_paths[i0:min(i0 + bs, hi)]:
            try:
                tensors.append(transform(Image.open(p).convert("RGB")))
                names.append(str(p))
            except Exception as e:  # noqa: BLE001 - unreadable images are skipped like the reference does
                errors += 1
                print(f"skipping image {p}: {e}")
        if not tensors:
            continue
        latent = vae_model.encode(torch.stack(tensors).pin_memory().to(device, non_blocking=True))
        flat = latent.reshape(latent.size(0), -1).cpu()
        for name, row in zip(names, flat):
            latent_data[name] = row.tolist()
    gathered = gather_to_rank0((latent_data, errors), rank, world)
    if own_group:
        import torch.distributed as dist

        dist.destroy_process_group()
    if rank != 0:
        return latent_data
    latent_data, errors = {}, 0
    for part, err in gathered:      # rank order = image-list order
        latent_data.update(part)
        errors += err
    print(f"done: {len(latent_data)} ok, {errors} failed, {len(image_paths)} total")
    out_path = Path(args.output_dir) / "latent_vectors.json"
    out_path.parent.mkdir(parents=True, exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(latent_data, f, indent=4)
    print(f"latent vectors saved to {out_path}")
    return latent_data


def build_parser():
    p = argparse.ArgumentParser(description="encode images with the VAE encoder and dump latent vectors (B200-native)")
    p.add_argument("--vae_checkpoint", type=str, required=True)
    p.add_argument("--vae_config_path", type=str, default=None)
    p.add_argument("--image_path", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="inference_output")
    p.add_argument("--resolution", type=int, default=1024)
    p.add_argument("--batch_size", type=int, default=8, help="images per GPU batch (addition of this implementation)")
    return p


def main(argv=None):
    return infer_and_save_latents(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
