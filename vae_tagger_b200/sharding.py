"""Batch sharding of an image stream over the GPUs of one box: every rank works on its own
images, no collective on the inference data path; the per-image results (a few bytes each) are gathered on the
host of rank 0.  Used by ``infer_full`` / ``infer_vae`` under ``torchrun`` and by ``tools/bulk_tag.py``."""
from __future__ import annotations

import os


def shard_range(num_items: int, rank: int, world: int):
    """Half-open index range of ``rank``: sizes differ by at most one, ranges tile [0, num_items)."""
    base, extra = divmod(num_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def image_cost(width: int, height: int) -> float:
    """Encoder work of one image in TFLOP: convs and projections are linear in the pixel count,
    the mid-block attention quadratic; p = pixels / 1024^2."""
    p = (width * height) / float(1024 * 1024)
    return 4.3329 * p + 0.54976 * p * p


def shard_by_cost(costs, rank: int, world: int):
    """Half-open index range of ``rank`` over items with the given per-item costs: contiguous spans of (nearly)
    equal total cost -- item i goes to the rank whose cost interval holds the midpoint of i's own interval, so
    the spans tile [0, n), depend only on (costs, world) and differ from the ideal by at most one item."""
    n = len(costs)
    if world <= 1:
        return 0, n
    total = float(sum(costs))
    if n == 0 or total <= 0.0:
        return shard_range(n, rank, world)
    bounds = [0] * (world + 1)
    bounds[world] = n
    acc, r = 0.0, 1
    for i, c in enumerate(costs):
        mid = acc + 0.5 * c
        while r < world and mid >= total * r / world:
            bounds[r] = i
            r += 1
        acc += c
    for k in range(r, world):
        bounds[k] = n
    return bounds[rank], bounds[rank + 1]


def dist_env():
    """(rank, world, local_rank) of this process as torchrun exports them; (0, 1, 0) for a plain run."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def init_host_group(world: int):
    """Process group for the host-side gather (gloo: python objects only; the GPUs exchange nothing).
    Returns True when this call created the group (the caller then destroys it)."""
    import torch.distributed as dist

    if world <= 1 or dist.is_initialized():
        return False
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group("gloo")
    return True


def gather_to_rank0(obj, rank: int, world: int):
    """List of every rank's ``obj`` (index = rank) on rank 0, ``None`` elsewhere; identity for one process."""
    if world <= 1:
        return [obj]
    import torch.distributed as dist

    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out
