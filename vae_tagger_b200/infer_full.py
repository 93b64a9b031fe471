"""Drop-in for the reference's ``infer_full.py`` (same CLI flags, same ``classification_results.json``).

What changes is *how* it runs: images are grouped into batches (and, with ``--use_bucketing``, into
aspect-ratio buckets), the FLUX VAE encoder + tag head run as sm_100a kernels, the descending sort
and the ``>= threshold`` count happen on the GPU, and each batch costs ONE device->host copy
instead of the reference's ``.item()`` per tag per image (infer_full.py:109-118).

    python -m vae_tagger_b200.infer_full --vae_checkpoint vae.safetensors --vae_config_path cfg.json \
        --decoder_checkpoint best_pytorch_model.bin --image_path imgs/ --tags_csv_path tags.csv
"""
from __future__ import annotations

import argparse
import json
import os
from pathlib import Path

import torch

from .diffusers_vae_loader import (DiffusersVAEWrapper, create_vae_from_config_file, get_diffusers_vae_config,
                                   load_diffusers_vae_from_config)
from .modules import (AspectRatioBucketing, ClassificationDecoder, create_attention_decoder, get_image_paths,
                      get_image_transform, get_vae_latent_info)


def load_models(args, device="cuda"):
    """VAE wrapper + decoder + tag names, with the reference's error behaviour (:16-71)."""
    import pandas as pd

    if args.vae_config_path and os.path.exists(args.vae_config_path):
        vae_model = create_vae_from_config_file(args.vae_config_path, args.vae_checkpoint)
    elif args.vae_checkpoint and os.path.exists(args.vae_checkpoint):
        vae_model = DiffusersVAEWrapper(load_diffusers_vae_from_config(get_diffusers_vae_config(), args.vae_checkpoint))
    else:
        raise RuntimeError("a VAE checkpoint or a VAE config file must be provided")
    vae_model.to(device).eval()
    info = get_vae_latent_info(args.resolution)
    tags_df = pd.read_csv(args.tags_csv_path)
    num_classes = len(tags_df)
    if args.use_attention:
        decoder = create_attention_decoder(
            info["latent_channels"], info["latent_height"], info["latent_width"], num_classes,
            attention_config={
                "use_spatial_attention": getattr(args, "use_spatial_attention", True),
                "use_self_attention": getattr(args, "use_self_attention", True),
                "use_cross_attention": getattr(args, "use_cross_attention", False),
                "attention_heads": getattr(args, "attention_heads", 8),
                "attention_dropout": getattr(args, "attention_dropout", 0.1),
            })
    else:
        decoder = ClassificationDecoder(info["latent_channels"], info["latent_height"], info["latent_width"],
                                        num_classes, use_adaptive_pooling=True)
    if not os.path.exists(args.decoder_checkpoint):
        raise RuntimeError(f"decoder checkpoint does not exist: {args.decoder_checkpoint}")
    try:
        decoder.load_state_dict(torch.load(args.decoder_checkpoint, map_location="cpu"), strict=False)
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(f"cannot load decoder checkpoint: {e}")
    decoder.to(device).eval()
    return vae_model, decoder, tags_df["name"].tolist()


def format_result(conf_row, idx_row, count, tag_names):
    """One image's JSON entry exactly as the reference builds it (:107-124) from HOST lists."""
    n = int(count)
    return {
        "predicted_tags": [{"tag": tag_names[int(i)], "confidence": float(f"{float(c):.4f}")}
                           for c, i in zip(conf_row[:n], idx_row[:n])],
        "total_tags_above_threshold": n,
        "max_confidence": float(f"{float(max(conf_row)):.4f}"),
        "avg_confidence_top5": float(f"{sum(float(c) for c in conf_row[:5]) / 5:.4f}"),
    }


@torch.no_grad()
def encode_and_tag(vae_model, decoder, pixel_values, threshold=0.5):
    """``decoder.get_confidence(vae_model.encode(x))`` + threshold count for a device batch (float [B,3,H,W] or
    uint8 [B,H,W,3]) as ONE native call (``vt_infer``): the tag head of every internal micro-batch runs right
    behind its encoder instead of trailing the whole batch.  Returns device tensors conf / idx / count / latent."""
    from .autoencoder_kl import AutoencoderKL

    vae = getattr(vae_model, "vae", vae_model)
    if not isinstance(vae, AutoencoderKL) or not decoder._use_native():
        latent = vae_model.encode(pixel_values)
        out = decoder.tag(latent, threshold=threshold)
        out["latent"] = latent
        return out
    ctx = vae._sync_native(vae._device_of(pixel_values))
    decoder._native_ctx(pixel_values.device)
    return ctx.infer(pixel_values, threshold=threshold, precision=vae._precision(), micro_batch=vae.micro_batch,
                     single_lane=vae.single_lane)


@torch.no_grad()
def classify_batch(vae_model, decoder, pixel_values, threshold):
    """encode -> get_confidence -> threshold for a [B,3,H,W] device batch; returns host lists."""
    out = encode_and_tag(vae_model, decoder, pixel_values, threshold)
    conf = out["conf"].cpu()   # the only device->host traffic of the batch
    idx = out["idx"].cpu()
    cnt = out["count"].cpu()
    return conf.tolist(), idx.tolist(), cnt.tolist()


def infer_and_classify(args):
    """The reference's ``infer_and_classify`` (infer_full.py:73-139).  Under ``torchrun`` (WORLD_SIZE > 1) the image
    list is sharded data-parallel over the GPUs of the box -- rank r takes a contiguous, cost-balanced span of the
    (bucket-grouped) list, nothing crosses NVLink -- the per-image results are gathered on rank 0's host and rank 0
    writes ONE ``classification_results.json`` identical to the single-GPU run."""
    from PIL import Image

    from .sharding import dist_env, gather_to_rank0, image_cost, init_host_group, shard_by_cost

    if not torch.cuda.is_available():
        raise RuntimeError("vae_tagger_b200 needs a CUDA device (B200); there is no CPU path")
    rank, world, local_rank = dist_env()
    if world > 1:
        local_rank %= torch.cuda.device_count()     # more ranks than GPUs (a test box): ranks share a device
        torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}" if world > 1 else "cuda"
    own_group = init_host_group(world)
    vae_model, decoder, tag_names = load_models(args, device)
    if not os.path.exists(args.image_path):
        raise FileNotFoundError(f"image path not found: {args.image_path}")
    image_paths = get_image_paths(args.image_path)
    if not image_paths:
        print("no image files found")
        return {}
    # group images: one group per target shape (a single square shape unless bucketing is on)
    groups = {}
    bucketing = AspectRatioBucketing(args.base_resolution, args.max_resolution, args.bucket_step) \
        if getattr(args, "use_bucketing", False) else None
    for p in image_paths:
        shape = bucketing.assign_bucket(str(p)) if bucketing else (args.resolution, args.resolution)
        groups.setdefault(shape, []).append(p)
    # the work list in processing order (group by group), and this rank's span of it
    work = [(shape, p) for shape, paths in groups.items() for p in paths]
    lo, hi = shard_by_cost([image_cost(w, h) for (w, h), _ in work], rank, world)
    my_groups = {}
    for shape, p in work[lo:hi]:
        my_groups.setdefault(shape, []).append(p)
    results, errors = {}, 0
    bs = max(1, getattr(args, "batch_size", 8))
    if getattr(args, "gpu_preprocess", False):
        # decode on the host, everything after that on the GPU: uint8 upload, crop + resize kernels
        # (bit-exact with PIL), ToTensor + Normalize fused into conv_in
        import numpy as np

        from .preprocess import BucketBatcher

        def decoded():
            nonlocal errors
            for _, p in work[lo:hi]:
                try:
                    yield str(p), np.array(Image.open(p).convert("RGB"))   # writable copy
                except Exception as e:  # noqa: BLE001 - the reference skips unreadable images (:130-132)
                    errors += 1
                    print(f"skipping image {p}: {e}")

        batcher = BucketBatcher(device, batch_size=bs, resolution=args.resolution, bucketing=bucketing)
        for _, names, batch in batcher.batches(decoded()):
            conf, idx, cnt = classify_batch(vae_model, decoder, batch, args.confidence_threshold)
            for name, c, i, n in zip(names, conf, idx, cnt):
                results[name] = format_result(c, i, n, tag_names)
        my_groups = {}
    for (w, h), paths in my_groups.items():
        tf = get_image_transform(args.resolution, bucketing is not None, (w, h) if bucketing else None)
        for i0 in range(0, len(paths), bs):
            tensors, names = [], []
            for p in paths[i0:i0 + bs]:
                try:
                    tensors.append(tf(Image.open(p).convert("RGB")))
                    names.append(str(p))
                except Exception as e:  # noqa: BLE001 - the reference skips unreadable images (:130-132)
                    errors += 1
                    print(f"skipping image {p}: {e}")
            if not tensors:
                continue
            batch = torch.stack(tensors).pin_memory().to(device, non_blocking=True)
            conf, idx, cnt = classify_batch(vae_model, decoder, batch, args.confidence_threshold)
            for name, c, i, n in zip(names, conf, idx, cnt):
                results[name] = format_result(c, i, n, tag_names)
    results, errors = merge_rank_results(gather_to_rank0((results, errors), rank, world), [str(p) for _, p in work])
    if own_group:
        import torch.distributed as dist

        dist.destroy_process_group()
    if rank != 0:
        return results
    print(f"done: {len(results)} ok, {errors} failed, {len(image_paths)} total")
    out_path = Path(args.output_dir) / "classification_results.json"
    out_path.parent.mkdir(parents=True, exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(results, f, indent=4, ensure_ascii=False)
    print(f"results saved to {out_path}")
    return results


def merge_rank_results(gathered, order):
    """Rank 0: merge the per-rank ``(results, errors)`` pairs into one dict in the processing order of the
    single-GPU run (``order`` = every image name, group by group); other ranks (``gathered is None``) get ({}, 0)."""
    if gathered is None:
        return {}, 0
    merged, errors = {}, 0
    for res, err in gathered:
        merged.update(res)
        errors += err
    return {name: merged[name] for name in order if name in merged}, errors


def build_parser():
    p = argparse.ArgumentParser(description="classify images with the FLUX VAE encoder + tag decoder (B200-native)")
    p.add_argument("--vae_checkpoint", type=str, required=True)
    p.add_argument("--vae_config_path", type=str, default=None)
    p.add_argument("--decoder_checkpoint", type=str, required=True)
    p.add_argument("--image_path", type=str, required=True)
    p.add_argument("--tags_csv_path", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="inference_output")
    p.add_argument("--resolution", type=int, default=1024)
    p.add_argument("--confidence_threshold", type=float, default=0.5)
    p.add_argument("--use_attention", action="store_true", default=True)
    p.add_argument("--no_attention", action="store_true")
    p.add_argument("--use_spatial_attention", action="store_true", default=True)
    p.add_argument("--use_self_attention", action="store_true", default=True)
    p.add_argument("--use_cross_attention", action="store_true")
    p.add_argument("--attention_heads", type=int, default=8)
    p.add_argument("--attention_dropout", type=float, default=0.1)
    p.add_argument("--model_checkpoint", type=str, default=None, help="(deprecated) parent of both checkpoints")
    # additions of this implementation (defaults keep the reference behaviour)
    p.add_argument("--batch_size", type=int, default=8, help="images per GPU batch")
    p.add_argument("--use_bucketing", action="store_true", help="group images by aspect-ratio bucket")
    p.add_argument("--base_resolution", type=int, default=512)
    p.add_argument("--max_resolution", type=int, default=1024)
    p.add_argument("--bucket_step", type=int, default=64)
    p.add_argument("--gpu_preprocess", action="store_true",
                   help="resize / crop / normalise on the GPU (bit-exact with the PIL transform) instead of the host")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.no_attention:
        args.use_attention = False
    if args.model_checkpoint and (not args.vae_checkpoint or not args.decoder_checkpoint):
        args.vae_checkpoint = args.model_checkpoint
        args.decoder_checkpoint = args.model_checkpoint
    return infer_and_classify(args)


if __name__ == "__main__":
    main()
