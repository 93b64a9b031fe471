"""torchrun check of the N>1 paths on real GPUs: decoder training step (NCCL gradient all-reduce, hidden
behind the next encoder forward) and parameter consistency across ranks.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402
from vae_tagger_b200 import modules as M  # noqa: E402
from vae_tagger_b200.improved_losses import FocalLoss  # noqa: E402
from vae_tagger_b200.train_decoder import DecoderTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
R, B, T = 512, 4, 1000
torch.manual_seed(0)
wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).to(dev).eval()
for p in wrap.parameters():
    p.requires_grad = False
torch.manual_seed(100 + rank)  # different init per rank: the trainer must broadcast rank 0's weights
dec = M.create_attention_decoder(16, R // 8, R // 8, T, attention_config={}).to(dev)
opt = torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-6)
tr = DecoderTrainer(wrap, dec, FocalLoss(1.0, 2.0), opt, None, max_grad_norm=1.0)
g = torch.Generator(device="cpu").manual_seed(7 + rank)
x = (torch.rand(B, 3, R, R, generator=g) * 2 - 1).to(dev)
y = (torch.rand(B, T, generator=g) < 0.1).float().to(dev)
losses = []
for it in range(6):
    if it == 2:
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    losses.append(tr.step(x, y))
tr.flush()
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / 4
chk = torch.stack([p.detach().double().sum() for p in dec.parameters()]).sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
if rank == 0:
    same = all(abs(c.item() - allc[0].item()) < 1e-9 * max(1.0, abs(allc[0].item())) for c in allc)
    print(f"ddp_check world={world}: losses {[round(l.item(), 5) for l in losses]}  step {dt * 1e3:.1f} ms "
          f"({world * B / dt:.1f} img/s at {R}^2)  params in sync: {same}")
    assert same and losses[-1] < losses[0]
dist.barrier()
dist.destroy_process_group()
