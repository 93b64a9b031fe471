"""CPU, world size 2, gloo: the N>1 host logic -- batch sharding of the inference stream and the
decoder-gradient exchange of the training step (one flat all-reduce, BN buffers from rank 0)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from vae_tagger_b200 import modules as M
from vae_tagger_b200.sharding import shard_range
from vae_tagger_b200.train_decoder import DecoderTrainer, cosine_schedule_with_warmup


def test_shard_ranges_cover_the_stream_without_overlap():
    for n in (0, 1, 7, 100_000):
        for g in (1, 2, 4, 8):
            spans = [shard_range(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cosine_schedule_matches_definition():
    opt = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=1.0)
    sched = cosine_schedule_with_warmup(opt, 2, 10)
    lrs = []
    for _ in range(10):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step(); sched.step()
    assert lrs[0] == 0.0 and abs(lrs[1] - 0.5) < 1e-9 and abs(lrs[2] - 1.0) < 1e-9
    assert abs(lrs[6] - 0.5) < 1e-9 and lrs[9] < 0.05


class _IdentityVAE(nn.Module):
    def encode(self, x):
        return x


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                 # deliberately different initial weights per rank
    dec = M.create_attention_decoder(16, 16, 16, 5, attention_config={})
    for m in dec.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    opt = torch.optim.SGD(dec.parameters(), lr=0.1)
    tr = DecoderTrainer(_IdentityVAE(), dec, nn.BCEWithLogitsLoss(), opt, None, max_grad_norm=0.0)
    start = {k: v.clone() for k, v in dec.state_dict().items()}
    g = torch.Generator().manual_seed(7 + rank)   # each rank has its own shard of the batch
    x = torch.randn(4, 16, 16, 16, generator=g)
    y = (torch.rand(4, 5, generator=g) < 0.3).float()
    tr.step(x, y)
    local_grad = tr.flat_grad.clone()             # before the exchange completes on this rank's view
    tr._pending.wait()
    summed = tr.flat_grad.clone()
    tr._pending = dist.all_reduce(torch.zeros(1), async_op=True)  # dummy handle so finish_update proceeds
    tr.flat_grad.copy_(summed)
    tr.finish_update()
    tr.flush()                                    # also lands the BatchNorm-buffer broadcast sent behind the step
    torch.save({"start": start, "end": dec.state_dict(), "summed": summed, "local": local_grad},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    # initial broadcast: both ranks start from rank 0's weights
    for k in r0["start"]:
        assert torch.equal(r0["start"][k], r1["start"][k]), k
    # the all-reduce sums the per-rank gradients identically on both ranks
    assert torch.allclose(r0["summed"], r1["summed"])
    # parameters stay in lock-step after the optimizer step; BatchNorm batch statistics are per rank (no SyncBN)
    # and the running buffers follow DDP's broadcast_buffers: every rank continues from rank 0's, sent right
    # behind the step that produced them
    for k in r0["end"]:
        if "running_" in k or "num_batches" in k:
            continue
        assert torch.allclose(r0["end"][k], r1["end"][k], atol=1e-7), k
    assert torch.equal(r0["end"]["feature_compress.1.running_mean"], r1["end"]["feature_compress.1.running_mean"])
    assert not torch.equal(r0["end"]["feature_compress.1.running_mean"], r0["start"]["feature_compress.1.running_mean"])


def _accum_data(rank, k):
    g = torch.Generator().manual_seed(50 + 10 * k + rank)
    return torch.randn(3, 16, 16, 16, generator=g) * 3, (torch.rand(3, 5, generator=g) < 0.3).float()


def _worker_accum(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(200 + rank)
    dec = M.create_attention_decoder(16, 16, 16, 5, attention_config={})
    for m in dec.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    opt = torch.optim.SGD(dec.parameters(), lr=0.1)
    tr = DecoderTrainer(_IdentityVAE(), dec, nn.BCEWithLogitsLoss(), opt, None, max_grad_norm=0.02,
                        gradient_accumulation_steps=2)
    start = {k: v.clone() for k, v in dec.state_dict().items()}
    for k in range(4):
        tr.step(*_accum_data(rank, k))
    tr.flush()
    torch.save({"start": start, "end": dec.state_dict()}, os.path.join(out_dir, f"a{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_accumulation_matches_the_reference_loop(tmp_path):
    """gradient_accumulation_steps = 2 on two ranks against a one-process emulation of the reference under DDP
    (train_decoder.py:186-203): every backward is averaged over the ranks, the accumulated gradient is clipped
    after every micro-step, the optimizer steps on every second one; BatchNorm statistics stay per rank."""
    world, port = 2, _free_port()
    mp.spawn(_worker_accum, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "a0.pt"), torch.load(tmp_path / "a1.pt")
    dec = M.create_attention_decoder(16, 16, 16, 5, attention_config={})
    for m in dec.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    dec.load_state_dict(r0["start"])
    dec.train()
    opt = torch.optim.SGD(dec.parameters(), lr=0.1)
    params = list(dec.parameters())
    acc = [torch.zeros_like(p) for p in params]
    loss_fn = nn.BCEWithLogitsLoss()
    for k in range(4):
        mean = [torch.zeros_like(p) for p in params]
        for rank in range(world):
            x, y = _accum_data(rank, k)
            grads = torch.autograd.grad(loss_fn(dec(x), y) / 2, params)
            for m_, g_ in zip(mean, grads):
                m_ += g_ / world
        for a_, m_ in zip(acc, mean):
            a_ += m_
        norm = torch.linalg.vector_norm(torch.stack([a_.norm() for a_ in acc]))
        coef = torch.clamp(0.02 / (norm + 1e-6), max=1.0)
        for a_ in acc:
            a_ *= coef
        if (k + 1) % 2 == 0:
            for p, a_ in zip(params, acc):
                p.grad = a_.clone()
            opt.step()
            for a_ in acc:
                a_.zero_()
    want = dec.state_dict()
    for k in want:
        if "running_" in k or "num_batches" in k:
            continue
        assert torch.allclose(r0["end"][k], r1["end"][k], atol=1e-7), k
        assert torch.allclose(r0["end"][k], want[k], atol=2e-6), (k, (r0["end"][k] - want[k]).abs().max())


# ----------------------------------------------------------------------------- sharded CLI (host logic)
def test_cost_balanced_shards_tile_the_work_list():
    from vae_tagger_b200.sharding import image_cost, shard_by_cost

    g = torch.Generator().manual_seed(0)
    shapes = [(512, 512), (1024, 1024), (832, 576), (640, 896)]
    for n in (0, 1, 3, 257):
        costs = [image_cost(*shapes[int(i)]) for i in torch.randint(0, 4, (n,), generator=g)]
        for world in (1, 2, 4, 8):
            spans = [shard_by_cost(costs, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            if n == 257:
                loads = [sum(costs[a:b]) for a, b in spans]
                assert max(loads) - min(loads) <= 2 * max(costs)       # within one item of the ideal on each side
    # 1024^2: 4.8826 TFLOP per image, attention quadratic in the pixel count
    assert abs(image_cost(1024, 1024) - 4.88266) < 1e-4 and abs(image_cost(512, 512) - 1.1176) < 1e-3


def _worker_shard_gather(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from vae_tagger_b200.infer_full import merge_rank_results
    from vae_tagger_b200.sharding import dist_env, gather_to_rank0, image_cost, init_host_group, shard_by_cost

    assert dist_env() == (rank, world, rank)
    own = init_host_group(world)
    assert own and dist.is_initialized()
    work = [((512 + 64 * (i % 5), 512), f"img{i:03d}.png") for i in range(23)]
    lo, hi = shard_by_cost([image_cost(w, h) for (w, h), _ in work], rank, world)
    results = {name: {"total_tags_above_threshold": int(name[3:6])} for _, name in work[lo:hi]}
    merged, errors = merge_rank_results(gather_to_rank0((results, rank), rank, world), [n for _, n in work])
    torch.save({"merged": merged, "errors": errors, "span": (lo, hi)}, os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_cli_gather(tmp_path):
    """infer_full under torchrun (reference call site infer_full.py:94-139, sharded per SURVEY 8e): every image is
    tagged by exactly one rank and rank 0 ends up with ONE result dict in the single-process order."""
    world, port = 2, _free_port()
    mp.spawn(_worker_shard_gather, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "s0.pt"), torch.load(tmp_path / "s1.pt")
    assert r0["span"][0] == 0 and r0["span"][1] == r1["span"][0] and r1["span"][1] == 23
    assert list(r0["merged"].keys()) == [f"img{i:03d}.png" for i in range(23)]
    assert all(v["total_tags_above_threshold"] == int(k[3:6]) for k, v in r0["merged"].items())
    assert r0["errors"] == 1 and r1["merged"] == {} and r1["errors"] == 0


def _worker_flat_allreduce(rank, world, port, out_dir):
    """train_full / train_vae: one flat all-reduce over every gradient of the step (mean over ranks), parameters
    without a gradient are skipped (a frozen or unused tensor must not desynchronise the flat layout)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vae_tagger_b200.train_full import _allreduce_grads
    torch.manual_seed(3)
    params = [nn.Parameter(torch.zeros(4, 3)), nn.Parameter(torch.zeros(5)), nn.Parameter(torch.zeros(2, 2))]
    g = torch.Generator().manual_seed(10 + rank)
    params[0].grad = torch.randn(4, 3, generator=g)
    params[2].grad = torch.randn(2, 2, generator=g)      # params[1] has no gradient on any rank
    local = [None if p.grad is None else p.grad.clone() for p in params]
    _allreduce_grads(params, world)
    torch.save({"local": local, "reduced": [None if p.grad is None else p.grad.clone() for p in params]},
               os.path.join(out_dir, f"f{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker_flat_allreduce, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "f0.pt"), torch.load(tmp_path / "f1.pt")
    for i in (0, 2):
        want = (r0["local"][i] + r1["local"][i]) / 2
        assert torch.allclose(r0["reduced"][i], want) and torch.equal(r0["reduced"][i], r1["reduced"][i])
    assert r0["reduced"][1] is None and r1["reduced"][1] is None
