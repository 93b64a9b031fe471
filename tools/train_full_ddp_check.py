"""Runs the drop-in train_full command line under torchrun on N GPUs with a tiny synthetic dataset (native
encoder backward + one flat gradient all-reduce per step) and checks its outputs; every rank also dumps a checksum of
its final encoder weights, which must agree across ranks (same start, all-reduced gradients, same optimizer).
    python tools/train_full_ddp_check.py [N]"""
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
for i in range(24):
    c = i % 3
    arr = torch.randint(0, 60, (64, 64, 3), generator=g, dtype=torch.uint8)
    arr[..., c] += 150
    path = os.path.join(d, f"im{i}.png")
    Image.fromarray(arr.numpy()).save(path)
    data[path] = names[c]
open(os.path.join(d, "data.json"), "w").write(json.dumps(data))
out = os.path.join(d, "out")
runner = os.path.join(d, "run.py")
open(runner, "w").write(
    "import os, sys, json, torch\n"
    f"sys.path.insert(0, {ROOT!r})\n"
    "from vae_tagger_b200 import train_full\n"
    "import vae_tagger_b200.train_full as tf\n"
    "keep = []\n"
    "orig = tf.create_vae_from_config_file\n"
    "def capture(*a, **k):\n"
    "    m = orig(*a, **k)\n"
    "    keep.append(m)\n"
    "    return m\n"
    "tf.create_vae_from_config_file = capture\n"
    "hist = train_full.main(sys.argv[1:])\n"
    "cs = sum(p.double().sum().item() for p in keep[0].vae.encoder.parameters())\n"
    "open(os.path.join(os.environ['VT_OUT'], f'rank{os.environ.get(\"RANK\", 0)}.json'), 'w').write(json.dumps({'checksum': cs}))\n")
env = dict(os.environ, VT_OUT=d)
cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
       "--master-port", "29541", runner, "--vae_checkpoint", os.path.join(d, "none.safetensors"),
       "--vae_config_path", os.path.join(d, "vae.json"), "--json_path", os.path.join(d, "data.json"), "--tags_csv_path",
       os.path.join(d, "tags.csv"), "--output_dir", out, "--resolution", "64", "--train_batch_size", "2", "--num_epochs",
       "2", "--num_workers", "0", "--lr_warmup_steps", "1", "--learning_rate", "2e-4", "--use_focal_loss", "--seed", "1"]
r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
print(r.stdout[-1500:])
if r.returncode != 0:
    print(r.stderr[-3000:])
    sys.exit(1)
hist = json.load(open(os.path.join(out, "training_history.json")))
sums = [json.load(open(os.path.join(d, f"rank{k}.json")))["checksum"] for k in range(n)]
assert len(hist["train_loss"]) == 2 and all(abs(s - sums[0]) <= 1e-9 * abs(sums[0]) for s in sums), (hist, sums)
print(f"train_full_ddp_check world={n}: train_loss {hist['train_loss']} encoder checksums {sums}: ok")
