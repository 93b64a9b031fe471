"""Golden sigmoid outputs / latents of the reference-side path on the BENCHED workload (run in the build
container only; ~30 min of CPU, resumable: per-image oracle latents are cached in /tmp):

    python tests/golden/make_tagset_golden.py

Path that produces them (all fp32, CPU): synthetic image -> ``oracle.encoder`` (restatement of diffusers'
``AutoencoderKL.encode``; diffusers itself is not installable here) -> ``DiffusersVAEWrapper.encode`` scale/shift
(diffusers_vae_loader.py:78-86) -> the REFERENCE's own ``AttentionClassificationDecoder`` (modules.py:358-475,
imported from /root/reference through the stub ``diffusers`` module of make_golden.py) -> sigmoid (infer_full.py
:100-118, modules.py:470-475).  Encoder weights: ``make_oracle_vae(seed=0)`` (random init, as the north star
states).  Heads:

  * ``T11`` / ``T1000``: the RANDOM-INIT state dicts of ``head_golden.pt`` (T = 11 is the reference's
    example_tags.csv vocabulary size, T = 1000 the benched one).  A random-init head is almost blind to its input
    (its LayerNorms see a bias-dominated vector): every image gets nearly the same probabilities.
  * ``T11_trained`` / ``T1000_trained``: the same architecture TRAINED here with the reference's own
    ``FocalLoss`` + AdamW + clip (train_decoder.py:186-203) for 300 steps on oracle latents of 160 structured
    images, against labels that are random half-spaces of the image-generation parameters.  Its probabilities
    spread over (0, 1) and react to the latent (8e-3 relative noise on the latent moves a sigmoid by up to 7e-3),
    which is what makes "identical tag sets" a test of the ENCODER's precision.  One trunk with 1011 outputs,
    split into the 11- and the 1000-tag head; the weights are rounded to fp16-representable values (stored as
    fp16 in ``trained_head.pt``) BEFORE the golden outputs are computed.

Image set (index order is the fixture's row order):
  * 112 at 256x256, 112 at 512x512, 32 at 1024x1024 (BASELINE configs[1]) -- even index: uniform noise
    (``synthetic_images``), odd index: ``structured_images``;
  * 8 reachable AspectRatioBucketing buckets (modules.py:188-222), one structured image each.
Stored: probabilities ``[N,T]`` fp32 for the four heads, image spec per row, and the fp16-rounded latents of the
first 8 images at 1024x1024, of 4 images each at 256 / 512 and of the 8 bucket images (fp16 storage error 2.8e-4
rms relative, far below the 1e-2 bar it is used against).
"""
import contextlib
import io
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import REF, install_stub_diffusers  # noqa: E402
from oracle.encoder import (make_oracle_vae, oracle_wrapper_encode, structured_images,  # noqa: E402
                            synthetic_images)

BUCKETS_HW = [(832, 576), (512, 1024), (960, 704), (1024, 576), (640, 896), (768, 1024), (1024, 960), (576, 512)]
QUIET = contextlib.redirect_stdout(io.StringIO())     # the reference's forward prints shapes


def image_specs():
    """(kind, H, W, seed) per fixture row."""
    specs = []
    for res, n in ((256, 112), (512, 112), (1024, 32)):
        for i in range(n):
            specs.append(("uniform" if i % 2 == 0 else "structured", res, res, 100000 * (res // 256) + i))
    for j, (h, w) in enumerate(BUCKETS_HW):
        specs.append(("structured", h, w, 900000 + j))
    return specs


def make_image(kind, h, w, seed):
    return (synthetic_images if kind == "uniform" else structured_images)(1, h, w, seed=seed)


def keep_latent(spec):
    kind, h, w, seed = spec
    if h == w == 1024:
        return (seed % 100000) < 8
    if h == w:
        return (seed % 100000) < 4
    return True


def train_head(ref_modules, ref_losses, vae):
    """The reference's head, trained with the reference's loss on oracle latents (see the module docstring).
    Returns the fp16-rounded state dict of the 1011-output model."""
    n, t_out = 160, 1011
    cache = "/tmp/tagset_train_latents.pt"
    if os.path.exists(cache):
        d = torch.load(cache)
        lat, par = d["lat"], d["par"]
    else:
        x, par = structured_images(n, 256, 256, seed=700000, return_params=True)
        with torch.no_grad():
            lat = torch.cat([oracle_wrapper_encode(vae, x[i:i + 8]) for i in range(0, n, 8)])
        torch.save({"lat": lat, "par": par}, cache)
    g = torch.Generator().manual_seed(11)
    feat = torch.cat([par, par[:, 2:5].abs(), par[:, 1:2] * par[:, 5:6]], 1)
    feat = (feat - feat.mean(0)) / feat.std(0)
    r = torch.randn(feat.shape[1], t_out, generator=g)
    b = torch.randn(t_out, generator=g) * 0.8 - 0.6
    y = ((feat @ r) / feat.shape[1] ** 0.5 * 1.5 + b > 0).float()
    torch.manual_seed(0)
    with QUIET:
        dec = ref_modules.create_attention_decoder(16, 32, 32, t_out, attention_config={})
    opt = torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-6)
    loss_fn = ref_losses.FocalLoss(alpha=1, gamma=2)
    dec.train()
    for step in range(300):
        idx = torch.randint(0, n, (32,), generator=g)
        with QUIET:
            loss = loss_fn(dec(lat[idx]), y[idx])
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(dec.parameters(), 1.0)
        opt.step()
        if step % 100 == 0:
            print(f"head training step {step}: focal loss {loss.item():.4f}", flush=True)
    return {k: (v.half().float() if v.is_floating_point() else v.clone()) for k, v in dec.state_dict().items()}


def split_head(sd, lo, hi):
    out = dict(sd)
    out["classifier.12.weight"] = sd["classifier.12.weight"][lo:hi].clone()
    out["classifier.12.bias"] = sd["classifier.12.bias"][lo:hi].clone()
    return out


def main():
    install_stub_diffusers()
    sys.path.insert(0, REF)
    import improved_losses as ref_losses  # the reference's own loss
    import modules as ref_modules  # the reference's own head

    golden = torch.load(os.path.join(HERE, "head_golden.pt"), map_location="cpu", weights_only=False)
    vae = make_oracle_vae(seed=0)
    trained = train_head(ref_modules, ref_losses, vae)
    torch.save({k: (v.half() if v.is_floating_point() else v) for k, v in trained.items()},
               os.path.join(HERE, "trained_head.pt"))
    state_dicts = {"T11_trained": split_head(trained, 0, 11), "T1000_trained": split_head(trained, 11, 1011)}
    for name, case in (("T11", "att_T11_64x64"), ("T1000", "att_T1000_16x16")):
        sd = dict(golden["attention_head_base"])
        sd.update(golden["attention_head"][case]["state_dict"])
        state_dicts[name] = sd
    heads = {}
    for name, sd in state_dicts.items():
        with QUIET:
            dec = ref_modules.create_attention_decoder(16, 128, 128, sd["classifier.12.bias"].numel(),
                                                       attention_config={}).eval()
        dec.load_state_dict(sd)
        heads[name] = dec

    specs = image_specs()
    cache_dir = "/tmp/tagset_golden_latents"
    os.makedirs(cache_dir, exist_ok=True)
    probs = {k: [] for k in heads}
    latents = {}
    t0 = time.time()
    for row, spec in enumerate(specs):
        f = os.path.join(cache_dir, f"row{row:03d}.pt")
        if os.path.exists(f):
            lat = torch.load(f)
        else:
            with torch.no_grad():
                lat = oracle_wrapper_encode(vae, make_image(*spec))
            torch.save(lat, f)
        with torch.no_grad(), QUIET:
            for k, d in heads.items():
                probs[k].append(torch.sigmoid(d(lat))[0].clone())
        if keep_latent(spec):
            latents[row] = lat[0].to(torch.float16)
        if row % 8 == 7 or row == len(specs) - 1:
            print(f"row {row + 1}/{len(specs)}  {time.time() - t0:.0f}s", flush=True)

    out = {"specs": specs, "latents": latents,
           "head_cases": {"T11": "att_T11_64x64", "T1000": "att_T1000_16x16"},
           "trained_split": {"T11_trained": (0, 11), "T1000_trained": (11, 1011)}}
    for k, v in probs.items():
        out["probs_" + k] = torch.stack(v)
    path = os.path.join(HERE, "tagset_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
