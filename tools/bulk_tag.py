"""BASELINE config 4: batch-sharded bulk tagging of a synthetic image stream at 1024^2, no collective.

Rank r of G tags the contiguous shard ``shard_range(N, r, G)`` of an N-image stream.  Images arrive as uint8 HWC
in pinned host memory (3 MB per image: what a decoder / data loader hands over), go through ``vt_infer_host``
(H2D per micro-batch under the previous micro-batch's kernels, encode, tag, D2H of the sorted confidences) and
only per-image tag counts / top tags stay on the host.  The host side is double buffered: while the GPU works
on batch k the next pinned batch is prepared.  Prints one JSON line (rank 0): aggregate images/s, wall clock,
max over ranks.

    python tools/bulk_tag.py [--images 2048] [--batch 32]
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port 29513 tools/bulk_tag.py --images 100000
"""
import argparse
import json
import os
import sys
import threading
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402
from vae_tagger_b200.sharding import shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2048, help="length of the whole stream (all ranks)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--tags", type=int, default=1000)
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches cycled through")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")  # host-side barrier / max only: the data path has no collective
    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).to(dev).eval()
    dec = M.create_attention_decoder(16, a.res // 8, a.res // 8, a.tags, attention_config={}).to(dev).eval()
    ctx = wrap.vae._sync_native(dev)
    dec._native_ctx(dev)
    lo, hi = shard_range(a.images, rank, world)
    g = torch.Generator().manual_seed(1000 + rank)
    pool = [torch.randint(0, 256, (a.batch, a.res, a.res, 3), generator=g, dtype=torch.uint8) for _ in range(a.pool)]
    bufs = [torch.empty(a.batch, a.res, a.res, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
    outs = [None, None]
    counts = torch.zeros(hi - lo, dtype=torch.int32)

    def fill(slot, k, n):  # stands in for decode + collate of the next batch
        bufs[slot][:n].copy_(pool[k % a.pool][:n])

    starts = list(range(lo, hi, a.batch))
    fill(0, 0, min(a.batch, hi - lo))
    ctx.infer_host(bufs[0], threshold=0.5)  # warm-up: workspace allocation, first-launch costs
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k, s0 in enumerate(starts):
        n = min(a.batch, hi - s0)
        slot = k & 1
        nxt = None
        if k + 1 < len(starts):
            nxt = threading.Thread(target=fill, args=(slot ^ 1, k + 1, min(a.batch, hi - starts[k + 1])))
            nxt.start()
        outs[slot] = ctx.infer_host(bufs[slot][:n], threshold=0.5, out=outs[slot] if n == a.batch else None)
        counts[s0 - lo:s0 - lo + n] = outs[slot]["count"][:n]
        if nxt is not None:
            nxt.join()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "images/s bulk tagging (uint8 host stream -> tags), end to end", "n_gpus": world,
                          "images": a.images, "batch": a.batch, "resolution": a.res, "tags": a.tags,
                          "value": round(a.images / t.item(), 2), "seconds": round(t.item(), 3),
                          "h2d_bytes_per_image": a.res * a.res * 3, "collectives_on_data_path": 0,
                          "shard_of_rank0": [lo, hi], "mean_tags_above_threshold": round(counts.float().mean().item(), 2)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
