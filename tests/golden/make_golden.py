"""Generate golden vectors from the reference's OWN code (run in the build container only).

    python tests/golden/make_golden.py

Imports ``modules.py`` / ``improved_losses.py`` / ``diffusers_vae_loader.py`` unchanged from
``/root/reference`` after installing a stub ``diffusers`` package in ``sys.modules`` (the real
one is not installed; those files only need the *name* ``AutoencoderKL`` at import time).
Writes small fixtures next to this script:

  head_golden.pt   state_dict + latent inputs + outputs of AttentionClassificationDecoder,
                   ClassificationDecoder, get_confidence, FocalLoss, AspectRatioBucketing,
                   get_vae_latent_info, DiffusersVAEWrapper.encode scale/shift.

The fixtures travel to the GPU box; ``/root/reference`` does not.
"""
import contextlib
import io
import os
import sys
import types

import torch

REF = os.environ.get("VT_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def install_stub_diffusers():
    class _StubAutoencoderKL(torch.nn.Module):  # never instantiated by the code we call
        def __init__(self, **kw):
            super().__init__()

    d = types.ModuleType("diffusers")
    dm = types.ModuleType("diffusers.models")
    do = types.ModuleType("diffusers.optimization")
    dm.AutoencoderKL = _StubAutoencoderKL
    do.get_scheduler = lambda *a, **k: None
    d.models, d.optimization = dm, do
    sys.modules.update({"diffusers": d, "diffusers.models": dm, "diffusers.optimization": do})


def main():
    install_stub_diffusers()
    sys.path.insert(0, REF)
    import diffusers_vae_loader as ref_loader  # noqa: E402
    import improved_losses as ref_losses  # noqa: E402
    import modules as ref_modules  # noqa: E402

    out = {}
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- attention head, T=11 (example_tags.csv) at 64x64 latent, and T=37 at 40x24 (ragged pool)
    cases = {}
    base_sd = None
    for name, (T, lh, lw, B, seed) in {
        "att_T11_64x64": (11, 64, 64, 3, 0),
        "att_T37_40x24": (37, 40, 24, 2, 1),
        "att_T1000_16x16": (1000, 16, 16, 2, 2),
    }.items():
        torch.manual_seed(seed)
        with quiet:
            dec = ref_modules.create_attention_decoder(16, lh, lw, T, attention_config={}).eval()
        # the head's parameters do not depend on the latent size: share everything but the
        # final T-way layer between cases so the fixture stays small
        if base_sd is not None:
            dec.load_state_dict({k: v for k, v in base_sd.items() if not k.startswith("classifier.12.")},
                                strict=False)
        # make BatchNorm running stats non-trivial so eval-mode BN is exercised
        with torch.no_grad():
            dec.feature_compress[1].running_mean.uniform_(-0.2, 0.2)
            dec.feature_compress[1].running_var.uniform_(0.5, 1.5)
            dec.feature_compress[1].weight.uniform_(0.5, 1.5)
            dec.feature_compress[1].bias.uniform_(-0.3, 0.3)
        lat = torch.randn(B, 16, lh, lw) * 0.5 + 0.1
        with torch.no_grad(), quiet:
            sa = dec.spatial_attention(lat)
            fc = dec.feature_compress(sa)
            at = dec.self_attention_post(fc)
            logits = dec(lat)
            conf, idx = dec.get_confidence(lat)
        full_sd = {k: v.clone() for k, v in dec.state_dict().items()}
        if base_sd is None:
            base_sd, stored = full_sd, full_sd
        else:
            stored = {k: v for k, v in full_sd.items()
                      if k.startswith("classifier.12.") or k.startswith("feature_compress.1.")}
        cases[name] = {
            "state_dict": stored,  # overrides on top of out["attention_head_base"]
            "latent": lat, "spatial": sa, "compressed": fc, "attended": at,
            "logits": logits, "conf": conf, "idx": idx,
        }
    out["attention_head"] = cases
    out["attention_head_base"] = base_sd

    # ---- plain head (--no_attention)
    torch.manual_seed(3)
    with quiet:
        pdec = ref_modules.ClassificationDecoder(16, 64, 64, 11).eval()
    lat = torch.randn(2, 16, 64, 64)
    with torch.no_grad(), quiet:
        out["plain_head"] = {
            "state_dict": {k: v.clone() for k, v in pdec.state_dict().items()},
            "latent": lat, "logits": pdec(lat),
        }

    # ---- parameter counts / key lists (SURVEY 8c known answers 2)
    with quiet:
        d1000 = ref_modules.create_attention_decoder(16, 128, 128, 1000, attention_config={})
    out["att_T1000_param_count"] = sum(p.numel() for p in d1000.parameters())
    out["att_keys"] = list(d1000.state_dict().keys())

    # ---- focal loss
    torch.manual_seed(4)
    x = torch.randn(5, 13) * 3
    y = (torch.rand(5, 13) < 0.2).float()
    fl = {}
    for a, g in ((1.0, 2.0), (0.25, 2.0), (1.0, 0.0), (0.5, 1.5)):
        xx = x.clone().requires_grad_(True)
        loss = ref_losses.FocalLoss(alpha=a, gamma=g)(xx, y)
        loss.backward()
        fl[(a, g)] = {"loss": loss.detach(), "grad": xx.grad.clone()}
    out["focal"] = {"logits": x, "targets": y, "cases": fl,
                    "at_zero": ref_losses.FocalLoss(1, 2)(torch.zeros(4, 7), torch.ones(4, 7))}

    # ---- bucketing / latent info
    arb = ref_modules.AspectRatioBucketing()
    ratios = {}
    for (w, h) in arb.buckets:  # first-in-sorted-order wins ties
        ratios.setdefault(w / h, (w, h))
    out["buckets"] = list(arb.buckets)
    out["reachable_buckets"] = sorted(ratios.values())
    out["latent_info_1024"] = ref_modules.get_vae_latent_info(1024)

    # ---- wrapper scale/shift through the reference's DiffusersVAEWrapper.encode
    class _FakeDist:
        def __init__(self, m):
            self._m = m

        def mode(self):
            return self._m

    class _FakeVAE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = types.SimpleNamespace(scaling_factor=0.3611, shift_factor=0.1159)

        def encode(self, x):
            return types.SimpleNamespace(latent_dist=_FakeDist(x))

    m = torch.randn(2, 16, 4, 4)
    out["wrapper"] = {"mean": m, "latent": ref_loader.DiffusersVAEWrapper(_FakeVAE()).encode(m)}
    out["vae_config"] = ref_loader.get_diffusers_vae_config()

    path = os.path.join(HERE, "head_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()


def dump_cli_flags():
    """CLI flag names of the reference entry points -> tests/golden/cli_flags.json"""
    import json
    import re

    out = {}
    for f in ("infer_full.py", "train_decoder.py", "infer_vae.py"):
        src = open(os.path.join(REF, f)).read()
        out[f] = sorted(set(re.findall(r'add_argument\(\s*"(--[a-z_0-9]+)"', src)))
    with open(os.path.join(HERE, "cli_flags.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    dump_cli_flags()
