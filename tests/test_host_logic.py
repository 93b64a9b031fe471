"""CPU: host-side logic of the drop-in surface (no GPU, no compute through the native library)."""
import json

import pytest
import torch

from oracle import encoder as OE
from vae_tagger_b200 import _native
from vae_tagger_b200 import diffusers_vae_loader as L
from vae_tagger_b200 import modules as M


def test_bucketing_matches_reference(golden):
    arb = M.AspectRatioBucketing()
    assert [tuple(b) for b in golden["buckets"]] == arb.buckets
    assert len(arb.buckets) == 81 and arb.buckets[0] == (512, 512) and arb.buckets[-1] == (1024, 1024)
    reach = {}
    for w, h in arb.buckets:
        reach.setdefault(w / h, (w, h))
    assert sorted(reach.values()) == [tuple(b) for b in golden["reachable_buckets"]]
    assert len(reach) == 67
    # first-in-sorted-order wins ties: a square image lands in (512, 512), never (1024, 1024)
    assert arb.bucket_for_size(1024, 1024) == (512, 512)
    assert arb.bucket_for_size(2000, 1000) == (1024, 512)


def test_latent_info_and_config(golden):
    assert M.get_vae_latent_info(1024) == golden["latent_info_1024"]
    assert M.get_vae_latent_info(1024)["total_dim"] == 262144
    assert L.get_diffusers_vae_config() == golden["vae_config"]


def test_head_state_dict_keys_match_reference(golden):
    dec = M.create_attention_decoder(16, 128, 128, 1000, attention_config={})
    assert list(dec.state_dict().keys()) == golden["att_keys"]
    assert sum(p.numel() for p in dec.parameters()) == golden["att_T1000_param_count"]
    plain = M.create_attention_decoder(16, 64, 64, 11, attention_config=None)
    assert list(plain.state_dict().keys()) == list(golden["plain_head"]["state_dict"].keys())
    assert isinstance(plain, M.ClassificationDecoder)


def test_vae_state_dict_keys_match_oracle_and_load():
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    oracle = OE.make_oracle_vae(0)
    assert sorted(vae.state_dict().keys()) == sorted(oracle.state_dict().keys())
    assert sum(p.numel() for p in vae.parameters()) == 34_274_208
    sd = dict(oracle.state_dict())
    # legacy checkpoints name the attention projections query/key/value/proj_attn
    for old, new in (("query", "to_q"), ("key", "to_k"), ("value", "to_v"), ("proj_attn", "to_out.0")):
        for leaf in ("weight", "bias"):
            sd[f"encoder.mid_block.attentions.0.{old}.{leaf}"] = sd.pop(f"encoder.mid_block.attentions.0.{new}.{leaf}")
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    assert torch.equal(vae.state_dict()["encoder.mid_block.attentions.0.to_q.weight"],
                       oracle.state_dict()["encoder.mid_block.attentions.0.to_q.weight"])
    assert vae.config.scaling_factor == 0.3611 and vae.config.shift_factor == 0.1159
    assert hasattr(vae.config, "scaling_factor")
    # a full FLUX checkpoint also carries decoder.* keys: they materialise the decoder half
    from oracle.decoder import make_oracle_decoder

    assert vae.decoder is None
    dec = make_oracle_decoder(1)
    full = dict(oracle.state_dict())
    full.update({"decoder." + k: v for k, v in dec.state_dict().items()})
    missing, unexpected = vae.load_state_dict(full, strict=False)
    assert not missing and not unexpected
    assert sorted(k for k in vae.state_dict() if k.startswith("decoder.")) == sorted("decoder." + k for k in dec.state_dict())
    assert torch.equal(vae.state_dict()["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"],
                       dec.state_dict()["up_blocks.2.resnets.0.conv_shortcut.weight"])
    assert sum(p.numel() for p in vae.decoder.parameters()) == 49_545_475
    assert sum(p.numel() for p in vae.parameters()) == 83_819_683      # SURVEY.md 8c known answer 1


def test_loader_file_roundtrip(tmp_path):
    cfg = tmp_path / "vae.json"
    cfg.write_text(json.dumps(L.get_diffusers_vae_config()))
    oracle = OE.make_oracle_vae(0)
    ck = tmp_path / "vae.pt"
    torch.save(oracle.state_dict(), ck)
    wrap = L.create_vae_from_config_file(str(cfg), str(ck))
    assert isinstance(wrap, L.DiffusersVAEWrapper)
    assert torch.equal(wrap.vae.state_dict()["encoder.conv_in.weight"], oracle.state_dict()["encoder.conv_in.weight"])
    # a missing checkpoint path silently keeps the random init (reference :37)
    wrap2 = L.create_vae_from_config_file(str(cfg), str(tmp_path / "missing.safetensors"))
    assert isinstance(wrap2.vae, torch.nn.Module)
    assert L.load_diffusers_vae_from_pretrained(str(tmp_path / "nope")) is None


def test_no_cpu_fallback():
    """The product path must fail loudly without the CUDA path (never route to a CPU implementation)."""
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).eval()
    with pytest.raises(_native.NativeError):
        wrap.encode(torch.zeros(1, 3, 64, 64))
    dec = M.create_attention_decoder(16, 8, 8, 11, attention_config={}).eval()
    with pytest.raises(_native.NativeError):
        dec(torch.zeros(1, 16, 8, 8))
    with pytest.raises(_native.NativeError):
        dec.get_confidence(torch.zeros(1, 16, 8, 8))
    with pytest.raises(_native.NativeError):      # the decoder half has no CPU path either
        wrap.vae.decode(torch.zeros(1, 16, 8, 8))
    with pytest.raises(_native.NativeError):
        wrap.decode(torch.zeros(1, 16, 8, 8))


def test_train_mode_head_is_differentiable_and_matches_oracle(golden):
    """train_decoder.py keeps a PyTorch graph in train() mode; with dropout off and BN in eval it must
    agree with the reference outputs."""
    c = golden["attention_head"]["att_T11_64x64"]
    sd = dict(golden["attention_head_base"]); sd.update(c["state_dict"])
    dec = M.create_attention_decoder(16, 64, 64, 11, attention_config={})
    dec.load_state_dict(sd)
    dec.train()
    for m in dec.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.BatchNorm2d)):
            m.eval()
    out = dec(c["latent"])
    assert torch.allclose(out, c["logits"], atol=1e-5)
    out.sum().backward()
    assert all(p.grad is not None for p in dec.parameters())


def test_package_does_not_import_oracle():
    import os
    import re

    root = os.path.dirname(os.path.abspath(M.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_cli_flags_cover_the_reference_entry_points():
    """Every flag of the reference's infer_full.py / train_decoder.py / infer_vae.py exists here (golden list
    parsed from the reference sources by tests/golden/make_golden.py)."""
    import os

    from vae_tagger_b200 import infer_full, infer_vae, train_decoder

    flags = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_flags.json")))
    for name, mod in (("infer_full.py", infer_full), ("train_decoder.py", train_decoder), ("infer_vae.py", infer_vae)):
        have = {a for act in mod.build_parser()._actions for a in act.option_strings}
        missing = [f for f in flags[name] if f not in have]
        assert not missing, (name, missing)
    # defaults that define behaviour
    a = infer_full.build_parser().parse_args(["--vae_checkpoint", "v", "--decoder_checkpoint", "d", "--image_path",
                                              "i", "--tags_csv_path", "t"])
    assert a.resolution == 1024 and a.confidence_threshold == 0.5 and a.use_attention and a.attention_heads == 8
    t = train_decoder.build_parser().parse_args(["--vae_checkpoint", "v", "--json_path", "j", "--tags_csv_path", "t"])
    assert t.learning_rate == 1e-3 and t.weight_decay == 1e-6 and t.lr_warmup_steps == 500 and t.max_grad_norm == 1.0
    assert t.focal_alpha == 1.0 and t.focal_gamma == 2.0 and t.gradient_accumulation_steps == 1 and t.seed == 42


def test_resize_coefficients_and_crop_box_match_the_oracle():
    """Host side of vt_resize_u8 (no GPU): the fixed-point filter taps and SmartResize's crop box computed
    by the library equal the oracle's restatement of Pillow / modules.py:149-172 bit for bit."""
    import numpy as np

    from oracle import resample as R
    from vae_tagger_b200 import _native

    rng = np.random.default_rng(5)
    pairs = [(4000, 1024), (97, 64), (64, 128), (513, 576), (1024, 1024), (1500, 512), (3, 7), (1, 1), (7, 3)]
    pairs += [(int(a), int(b)) for a, b in zip(rng.integers(1, 5000, 40), rng.integers(1, 1100, 40))]
    for a, b in pairs:
        for f in (R.LANCZOS, R.BILINEAR):
            ks, bd, kk = _native.resize_coefficients(a, b, f)
            ks2, bd2, kk2 = R.precompute_coeffs(a, b, f)
            assert ks == ks2 and np.array_equal(bd, bd2) and np.array_equal(kk, kk2), (a, b, f)
            assert (kk.sum(1) - (1 << 22)).__abs__().max() <= ks      # taps sum to one in fixed point
    from vae_tagger_b200.modules import AspectRatioBucketing

    arb = AspectRatioBucketing()
    for _ in range(200):
        ow, oh = int(rng.integers(64, 6000)), int(rng.integers(64, 6000))
        tw, th = arb.bucket_for_size(ow, oh)
        box = _native.smart_crop_box(ow, oh, tw, th)
        assert box == R.smart_crop_box(ow, oh, tw, th)
        l, t, r, b = box
        assert 0 <= l < r <= ow and 0 <= t < b <= oh


def test_triplet_dataset_mining(tmp_path):
    """``TaggedImageDataset(triplets=True)`` yields the keys train_full.py consumes (reference modules.py:628-648) and
    its mining follows the reference's rules: a positive shares a tag with the anchor whenever one exists among the
    other images, a negative shares none."""
    import json as _json
    import random

    from PIL import Image

    from vae_tagger_b200 import modules as M
    names = ["a", "b", "c"]
    (tmp_path / "tags.csv").write_text("name\n" + "\n".join(names) + "\n")
    data = {}
    for i in range(9):
        p = tmp_path / f"im{i}.png"
        Image.new("RGB", (40, 32), (i * 20, 0, 0)).save(p)
        data[str(p)] = ("a, b" if i % 3 == 0 else names[i % 3]) + (":0.5" if i == 4 else "")
    (tmp_path / "d.json").write_text(_json.dumps(data))
    ds = M.TaggedImageDataset(str(tmp_path / "d.json"), str(tmp_path / "tags.csv"), M.get_image_transform(32), triplets=True)
    random.seed(0)
    for idx in range(len(ds)):
        it = ds[idx]
        assert set(it) == {"pixel_values", "labels", "anchor", "positive", "negative", "positive_labels", "negative_labels"}
        assert it["anchor"].shape == it["positive"].shape == it["negative"].shape == (3, 32, 32)
        assert (it["positive_labels"] * it["labels"]).sum() > 0          # every tag has other carriers here
        assert (it["negative_labels"] * it["labels"]).sum() == 0
    plain = M.TaggedImageDataset(str(tmp_path / "d.json"), str(tmp_path / "tags.csv"), M.get_image_transform(32))
    assert set(plain[0]) == {"pixel_values", "labels"}
    assert plain.image_labels[str(tmp_path / "im4.png")][1].item() == 0.5


def test_save_pretrained_round_trip(tmp_path):
    from vae_tagger_b200 import diffusers_vae_loader as L
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    vae.save_pretrained(str(tmp_path / "v"))
    back = L.load_diffusers_vae_from_pretrained(str(tmp_path / "v"))
    assert back is not None and back.config.scaling_factor == 0.3611
    a, b = vae.state_dict(), back.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
