"""Runs the drop-in train_decoder command line under torchrun on N GPUs with a tiny synthetic dataset and checks
its outputs (rank 0 writes best_pytorch_model.bin + training_history.json):
    python tools/train_cli_ddp_check.py [N]"""
import json
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vae_tagger_b200 import diffusers_vae_loader as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = tempfile.mkdtemp()
from PIL import Image  # noqa: E402

open(os.path.join(d, "vae.json"), "w").write(json.dumps(L.get_diffusers_vae_config()))
names = ["red", "green", "blue"]
open(os.path.join(d, "tags.csv"), "w").write("name\n" + "\n".join(names) + "\n")
g = torch.Generator().manual_seed(3)
data = {}
for i in range(40):
    c = i % 3
    arr = torch.randint(0, 60, (64, 64, 3), generator=g, dtype=torch.uint8)
    arr[..., c] += 150
    path = os.path.join(d, f"im{i}.png")
    Image.fromarray(arr.numpy()).save(path)
    data[path] = names[c]
open(os.path.join(d, "data.json"), "w").write(json.dumps(data))
out = os.path.join(d, "out")
cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
       "--master-port", "29533", "-m", "vae_tagger_b200.train_decoder", "--vae_checkpoint", os.path.join(d, "none.safetensors"),
       "--vae_config_path", os.path.join(d, "vae.json"), "--json_path", os.path.join(d, "data.json"), "--tags_csv_path",
       os.path.join(d, "tags.csv"), "--output_dir", out, "--resolution", "64", "--train_batch_size", "4", "--num_epochs",
       "3", "--num_workers", "0", "--lr_warmup_steps", "1", "--learning_rate", "3e-3", "--use_focal_loss", "--seed", "1"]
r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
print(r.stdout[-1500:])
if r.returncode != 0:
    print(r.stderr[-3000:])
    sys.exit(1)
hist = json.load(open(os.path.join(out, "training_history.json")))
sd = torch.load(os.path.join(out, "best_pytorch_model.bin"), map_location="cpu")
assert len(hist["train_loss"]) == 3 and hist["train_loss"][-1] < hist["train_loss"][0], hist
print(f"train_cli_ddp_check world={n}: train_loss {hist['train_loss']} val_loss {hist['val_loss']} tensors {len(sd)}: ok")
