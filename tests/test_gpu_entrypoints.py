"""GPU: the reference-facing entry points end to end (infer_full CLI, aspect-ratio bucket shapes,
the train_decoder step) against the oracle."""
import json
import os

import pytest
import torch

from oracle import head as OH
from oracle.encoder import make_oracle_vae, oracle_wrapper_encode, synthetic_images
from vae_tagger_b200 import diffusers_vae_loader as L
from vae_tagger_b200 import modules as M

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_infer_full_cli_matches_oracle(tmp_path, golden, monkeypatch):
    from PIL import Image
    from safetensors.torch import save_file

    from vae_tagger_b200 import infer_full

    res = 64
    oracle = make_oracle_vae(0)
    save_file({k: v.contiguous() for k, v in oracle.state_dict().items()}, str(tmp_path / "vae.safetensors"))
    (tmp_path / "vae.json").write_text(json.dumps(L.get_diffusers_vae_config()))
    sd = dict(golden["attention_head_base"]); sd.update(golden["attention_head"]["att_T11_64x64"]["state_dict"])
    torch.save(sd, tmp_path / "decoder.bin")
    names = [f"tag{i}" for i in range(11)]
    (tmp_path / "tags.csv").write_text("name\n" + "\n".join(names) + "\n")
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    g = torch.Generator().manual_seed(0)
    for i in range(5):
        arr = torch.randint(0, 256, (48 + 8 * i, 80, 3), generator=g, dtype=torch.uint8).numpy()
        Image.fromarray(arr).save(img_dir / f"im{i}.png")
    (img_dir / "broken.png").write_bytes(b"not an image")  # skipped like the reference does (:130-132)

    monkeypatch.setenv("VT_B200_PRECISION", "fp32")
    out = infer_full.main(["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path",
                           str(tmp_path / "vae.json"), "--decoder_checkpoint", str(tmp_path / "decoder.bin"),
                           "--image_path", str(img_dir), "--tags_csv_path", str(tmp_path / "tags.csv"),
                           "--output_dir", str(tmp_path / "out"), "--resolution", str(res), "--batch_size", "2"])
    saved = json.loads((tmp_path / "out" / "classification_results.json").read_text())
    assert saved == out and len(saved) == 5
    tf = M.get_image_transform(res)
    for path, entry in saved.items():
        x = tf(Image.open(path).convert("RGB")).unsqueeze(0)
        with torch.no_grad():
            conf, idx = OH.get_confidence(OH.attention_decoder_logits(sd, oracle_wrapper_encode(oracle, x)))
        want = OH.threshold_tags(conf[0], idx[0], 0.5)
        assert entry["total_tags_above_threshold"] == want["total_tags_above_threshold"]
        assert [t["tag"] for t in entry["predicted_tags"]] == [names[i] for i, _ in want["predicted"]]
        for t, (_, c) in zip(entry["predicted_tags"], want["predicted"]):
            assert abs(t["confidence"] - c) <= 2e-4
        assert abs(entry["max_confidence"] - want["max_confidence"]) <= 2e-4
        assert abs(entry["avg_confidence_top5"] - want["avg_confidence_top5"]) <= 2e-4
    # --gpu_preprocess: same pixels (the resize kernels are bit-exact with PIL), hence the same JSON
    base = ["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path", str(tmp_path / "vae.json"),
            "--decoder_checkpoint", str(tmp_path / "decoder.bin"), "--image_path", str(img_dir), "--tags_csv_path",
            str(tmp_path / "tags.csv"), "--resolution", str(res), "--batch_size", "2"]

    def same(a, b):
        assert a.keys() == b.keys()
        for k in a:
            assert a[k]["total_tags_above_threshold"] == b[k]["total_tags_above_threshold"]
            assert [t["tag"] for t in a[k]["predicted_tags"]] == [t["tag"] for t in b[k]["predicted_tags"]]
            for ta, tb in zip(a[k]["predicted_tags"], b[k]["predicted_tags"]):
                assert abs(ta["confidence"] - tb["confidence"]) <= 1e-4
            assert abs(a[k]["max_confidence"] - b[k]["max_confidence"]) <= 1e-4

    same(infer_full.main(base + ["--output_dir", str(tmp_path / "out_gpu"), "--gpu_preprocess"]), saved)
    bucket = ["--use_bucketing", "--base_resolution", "64", "--max_resolution", "128", "--bucket_step", "32"]
    host = infer_full.main(base + bucket + ["--output_dir", str(tmp_path / "out_b")])
    assert len(host) == 5
    same(infer_full.main(base + bucket + ["--output_dir", str(tmp_path / "out_bg"), "--gpu_preprocess"]), host)
    with pytest.raises(RuntimeError):
        infer_full.main(["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--decoder_checkpoint",
                         str(tmp_path / "missing.bin"), "--image_path", str(img_dir), "--tags_csv_path",
                         str(tmp_path / "tags.csv")])


@pytest.mark.parametrize("W,H", [(512, 576), (640, 512)])
def test_bucket_shapes_bf16(W, H):
    """Non-square aspect-ratio buckets (BASELINE config 3): partial tiles at every level (576/8 = 72 is
    not a multiple of the 16-pixel tile)."""
    assert (W, H) in M.AspectRatioBucketing().buckets
    oracle = make_oracle_vae(0)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    vae.load_state_dict(oracle.state_dict())
    wrap = L.DiffusersVAEWrapper(vae).cuda().eval()
    x = synthetic_images(1, H, W)
    with torch.no_grad():
        ref = oracle_wrapper_encode(oracle, x)
    got = wrap.encode(x.cuda()).cpu()
    assert got.shape == (1, 16, H // 8, W // 8)
    assert rel(got, ref) <= 1e-2, rel(got, ref)


def test_focal_loss_module_autograd(golden):
    from vae_tagger_b200.improved_losses import FocalLoss

    f = golden["focal"]
    for (a, g), want in f["cases"].items():
        x = f["logits"].cuda().requires_grad_(True)
        loss = FocalLoss(alpha=a, gamma=g)(x, f["targets"].cuda())
        loss.backward()
        assert abs(loss.item() - want["loss"].item()) < 1e-6
        assert rel(x.grad.cpu(), want["grad"]) < 1e-5


def test_train_decoder_step_single_rank():
    """Config 5 on one rank: frozen native encoder -> head (train mode) -> fused focal loss -> AdamW."""
    from vae_tagger_b200.improved_losses import FocalLoss
    from vae_tagger_b200.train_decoder import DecoderTrainer

    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
    for p in wrap.parameters():
        p.requires_grad = False
    dec = M.create_attention_decoder(16, 8, 8, 11, attention_config={}).cuda()
    opt = torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-6)
    tr = DecoderTrainer(wrap, dec, FocalLoss(1.0, 2.0), opt, None, max_grad_norm=1.0)
    x = synthetic_images(8, 64, 64).cuda()
    y = (torch.rand(8, 11, generator=torch.Generator().manual_seed(1)) < 0.1).float().cuda()
    losses = [tr.step(x, y).item() for _ in range(12)]
    tr.flush()
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0], losses
    # the head in eval mode (native kernels) agrees with its own training graph in eval semantics
    dec.eval()
    lat = wrap.encode(x)
    native = dec(lat)
    dec.train()
    for m in dec.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.BatchNorm2d)):
            m.eval()
    with torch.enable_grad():
        graph = dec(lat)
    assert rel(native.cpu(), graph.detach().cpu()) < 2e-5


def test_fused_infer_call_equals_encode_then_tag(golden):
    """vt_infer (head per micro-batch behind its encoder) == DiffusersVAEWrapper.encode followed by decoder.tag,
    on float and uint8 inputs, with and without internal micro-batching."""
    from vae_tagger_b200.infer_full import encode_and_tag

    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
    sd = dict(golden["attention_head_base"]); sd.update(golden["attention_head"]["att_T11_64x64"]["state_dict"])
    dec = M.create_attention_decoder(16, 8, 16, 11, attention_config={})
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    x = synthetic_images(5, 64, 128).cuda()
    xu = ((x.permute(0, 2, 3, 1) * 0.5 + 0.5) * 255).round().clamp(0, 255).to(torch.uint8).contiguous()
    wrap.vae.precision = "fp32"
    for inp in (x, xu):
        for mb in (0, 2):
            wrap.vae.micro_batch = mb
            lat = wrap.encode(inp)
            want = dec.tag(lat, threshold=0.5)
            got = encode_and_tag(wrap, dec, inp, threshold=0.5)
            assert rel(got["latent"], lat) <= 1e-6
            assert (got["conf"] - want["conf"]).abs().max().item() <= 1e-6
            assert torch.equal(got["count"], want["count"])
            gap = (want["conf"][:, :-1] - want["conf"][:, 1:]).abs().min().item()
            if gap > 1e-5:
                assert torch.equal(got["idx"], want["idx"])
    wrap.vae.micro_batch = 0
    wrap.vae.precision = "bf16"
    host = x.cpu().pin_memory()
    ctx = wrap.vae._sync_native(x.device)
    out = ctx.infer_host(host, threshold=0.5)
    want = encode_and_tag(wrap, dec, x, threshold=0.5)
    assert torch.equal(out["conf"], want["conf"].cpu())      # two 16-bit runs: bit-identical
    assert torch.equal(out["idx"], want["idx"].cpu())


def test_config1_512_batch1_fp32(golden):
    """BASELINE configs[0] -- the reference's own CPU-runnable case: FLUX VAE encoder (random init) + 8-head
    attention tagger, 512x512, batch 1, fp32 -- through encode + tag, against the oracle run on the CPU:
    latent rel-L2 <= 1e-4 (north star, fp32 mode), max |delta sigmoid| <= 1e-4, identical tags at 0.5."""
    oracle = make_oracle_vae(0)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    vae.load_state_dict(oracle.state_dict(), strict=False)
    wrap = L.DiffusersVAEWrapper(vae).cuda().eval()
    wrap.vae.precision = "fp32"
    sd = dict(golden["attention_head_base"]); sd.update(golden["attention_head"]["att_T11_64x64"]["state_dict"])
    dec = M.create_attention_decoder(16, 64, 64, 11, attention_config={})
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    x = synthetic_images(1, 512, 512)
    with torch.no_grad():
        ref_lat = oracle_wrapper_encode(oracle, x)
        ref_p = torch.sigmoid(OH.attention_decoder_logits(sd, ref_lat))
    lat = wrap.encode(x.cuda())
    assert lat.shape == (1, 16, 64, 64) and rel(lat.cpu(), ref_lat) <= 1e-4, rel(lat.cpu(), ref_lat)
    conf, idx = dec.get_confidence(lat)
    p = torch.sigmoid(dec(lat)).cpu()
    assert (p - ref_p).abs().max().item() <= 1e-4
    assert torch.equal(p >= 0.5, ref_p >= 0.5)
    assert torch.equal(idx.cpu()[0], torch.sort(ref_p[0], descending=True).indices) or \
        (ref_p[0].sort(descending=True).values.diff().abs().min().item() < 1e-5)


def test_infer_vae_cli_matches_oracle(tmp_path, monkeypatch):
    """infer_vae.py (reference :31-81): latent_vectors.json = flattened mode()*scale+shift per image, against the
    oracle on the same host-transformed pixels (fp32 mode, rel-L2 <= 1e-4)."""
    from PIL import Image
    from safetensors.torch import save_file

    from vae_tagger_b200 import infer_vae

    oracle = make_oracle_vae(0)
    save_file({k: v.contiguous() for k, v in oracle.state_dict().items()}, str(tmp_path / "vae.safetensors"))
    (tmp_path / "vae.json").write_text(json.dumps(L.get_diffusers_vae_config()))
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    g = torch.Generator().manual_seed(2)
    for i in range(3):
        arr = torch.randint(0, 256, (40 + 16 * i, 72, 3), generator=g, dtype=torch.uint8).numpy()
        Image.fromarray(arr).save(img_dir / f"im{i}.png")
    monkeypatch.setenv("VT_B200_PRECISION", "fp32")
    out = infer_vae.main(["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path",
                          str(tmp_path / "vae.json"), "--image_path", str(img_dir), "--output_dir",
                          str(tmp_path / "out"), "--resolution", "64", "--batch_size", "2"])
    saved = json.loads((tmp_path / "out" / "latent_vectors.json").read_text())
    assert saved.keys() == out.keys() and len(saved) == 3
    tf = M.get_image_transform(64)
    for path, vec in saved.items():
        with torch.no_grad():
            want = oracle_wrapper_encode(oracle, tf(Image.open(path).convert("RGB")).unsqueeze(0)).reshape(-1)
        got = torch.tensor(vec)
        assert got.numel() == 16 * 8 * 8 and rel(got, want) <= 1e-4


def test_empty_batch_is_an_empty_result(golden):
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
    dec = M.create_attention_decoder(16, 8, 8, 11, attention_config={}).cuda().eval()
    lat = wrap.encode(torch.empty(0, 3, 64, 64, device="cuda"))
    assert lat.shape == (0, 16, 8, 8)
    conf, idx = dec.get_confidence(lat)
    assert conf.shape == (0, 11) and idx.shape == (0, 11) and idx.dtype == torch.int64
    assert dec(lat).shape == (0, 11)
    assert wrap.decode(lat).shape == (0, 3, 64, 64)


@pytest.mark.parametrize("extra", [["--use_focal_loss"], ["--use_class_balanced", "--gradient_accumulation_steps", "2"],
                                   ["--no_attention"],
                                   ["--use_bucketing", "--base_resolution", "64", "--max_resolution", "128",
                                    "--bucket_step", "32", "--train_batch_size", "1"],
                                   ["--use_focal_loss", "--gpu_preprocess"],
                                   ["--use_bucketing", "--base_resolution", "64", "--max_resolution", "128",
                                    "--bucket_step", "32", "--train_batch_size", "1", "--gpu_preprocess",
                                    "--num_workers", "2"]])
def test_train_decoder_cli_end_to_end(tmp_path, extra):
    """train_decoder.py's command line on a tiny dataset: JSON prompts + tags.csv + image files -> two epochs ->
    best_pytorch_model.bin (reference state-dict keys) + training_history.json, through the native training step."""
    from PIL import Image
    from safetensors.torch import save_file

    from vae_tagger_b200 import train_decoder

    oracle = make_oracle_vae(0)
    save_file({k: v.contiguous() for k, v in oracle.state_dict().items()}, str(tmp_path / "vae.safetensors"))
    (tmp_path / "vae.json").write_text(json.dumps(L.get_diffusers_vae_config()))
    names = ["red", "green", "blue", "dark"]
    (tmp_path / "tags.csv").write_text("name\n" + "\n".join(names) + "\n")
    g = torch.Generator().manual_seed(3)
    data = {}
    for i in range(20):
        c = i % 3
        arr = torch.randint(0, 60, (72, 80, 3), generator=g, dtype=torch.uint8)
        arr[..., c] += 150                                   # the tag is the dominant colour channel
        path = tmp_path / f"im{i}.png"
        Image.fromarray(arr.numpy()).save(path)
        data[str(path)] = f"{names[c]}:1.0, dark:0.5" if i % 5 == 0 else names[c]
    (tmp_path / "data.json").write_text(json.dumps(data))
    out_dir = tmp_path / "out"
    hist = train_decoder.main(["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path",
                               str(tmp_path / "vae.json"), "--json_path", str(tmp_path / "data.json"),
                               "--tags_csv_path", str(tmp_path / "tags.csv"), "--output_dir", str(out_dir),
                               "--resolution", "64", "--train_batch_size", "4", "--num_epochs", "3",
                               "--num_workers", "0", "--lr_warmup_steps", "1", "--learning_rate", "3e-3",
                               "--logging_steps", "2"] + extra)
    assert len(hist["train_loss"]) == 3 and all(torch.isfinite(torch.tensor(hist["train_loss"] + hist["val_loss"])))
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    saved = json.loads((out_dir / "training_history.json").read_text())
    assert saved == hist
    sd = torch.load(out_dir / "best_pytorch_model.bin", map_location="cpu")
    if "--no_attention" in extra:
        ref = M.ClassificationDecoder(16, 8, 8, 4)
    else:
        ref = M.create_attention_decoder(16, 8, 8, 4, attention_config={})
    assert list(sd.keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(sd)


def test_infer_full_sharded_under_torchrun(tmp_path, golden):
    """``torchrun --nproc-per-node 2 -m vae_tagger_b200.infer_full`` (north star: the image batch is sharded
    data-parallel, no collective on the inference path): one classification_results.json on rank 0, equal to
    the single-process run entry by entry and in the same order.  With fewer GPUs than ranks the ranks share
    cuda:0 -- the sharding / gather logic is the same.  Same for infer_vae's latent_vectors.json."""
    import socket
    import subprocess
    import sys

    from PIL import Image
    from safetensors.torch import save_file

    from vae_tagger_b200 import infer_full, infer_vae

    oracle = make_oracle_vae(0)
    save_file({k: v.contiguous() for k, v in oracle.state_dict().items()}, str(tmp_path / "vae.safetensors"))
    (tmp_path / "vae.json").write_text(json.dumps(L.get_diffusers_vae_config()))
    sd = dict(golden["attention_head_base"]); sd.update(golden["attention_head"]["att_T11_64x64"]["state_dict"])
    torch.save(sd, tmp_path / "decoder.bin")
    (tmp_path / "tags.csv").write_text("name\n" + "\n".join(f"tag{i}" for i in range(11)) + "\n")
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    g = torch.Generator().manual_seed(1)
    for i in range(7):
        arr = torch.randint(0, 256, (64 + 16 * (i % 3), 96, 3), generator=g, dtype=torch.uint8).numpy()
        Image.fromarray(arr).save(img_dir / f"im{i}.png")
    (img_dir / "broken.png").write_bytes(b"not an image")
    base = ["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path", str(tmp_path / "vae.json"),
            "--image_path", str(img_dir), "--resolution", "64", "--batch_size", "2"]
    tag_args = ["--decoder_checkpoint", str(tmp_path / "decoder.bin"), "--tags_csv_path", str(tmp_path / "tags.csv"),
                "--use_bucketing", "--base_resolution", "64", "--max_resolution", "128", "--bucket_step", "32"]
    single = infer_full.main(base + tag_args + ["--output_dir", str(tmp_path / "one")])
    single_lat = infer_vae.main(base + ["--output_dir", str(tmp_path / "one")])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    run = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), "-m"]
    logs = []
    for mod, extra in (("vae_tagger_b200.infer_full", tag_args), ("vae_tagger_b200.infer_vae", [])):
        out = subprocess.run(run + [mod] + base + extra + ["--output_dir", str(tmp_path / "two")], env=env, cwd=root,
                             capture_output=True, text=True, timeout=900)
        logs.append(out.stdout[-3000:] + out.stderr[-3000:])
        assert out.returncode == 0, logs[-1]
    sharded = json.loads((tmp_path / "two" / "classification_results.json").read_text())
    assert list(sharded.keys()) == list(single.keys()) and len(sharded) == 7, (list(sharded.keys()), logs[0])
    assert sharded == single       # bf16 mode is batch-invariant: bit-identical probabilities, hence identical JSON
    sharded_lat = json.loads((tmp_path / "two" / "latent_vectors.json").read_text())
    assert list(sharded_lat.keys()) == list(single_lat.keys()) and sharded_lat == single_lat


@pytest.mark.parametrize("extra", [["--use_focal_loss"], ["--use_full_loss", "--use_adaptive_weights", "--similarity_type", "euclidean"],
                                   ["--mixed_precision", "no", "--gradient_accumulation_steps", "2"]])
def test_train_full_cli_end_to_end(tmp_path, extra):
    """train_full.py's command line (reference step train_full.py:201-256): three native encoder training
    forwards per step, triplet + classification loss, native encoder backward, AdamW over the encoder AND the head.
    Two epochs on a tiny dataset: losses finite and falling, the fine-tuned VAE is saved in the diffusers layout and
    its encoder weights have moved."""
    from PIL import Image
    from safetensors.torch import load_file, save_file

    from vae_tagger_b200 import train_full

    oracle = make_oracle_vae(0)
    save_file({k: v.contiguous() for k, v in oracle.state_dict().items()}, str(tmp_path / "vae.safetensors"))
    (tmp_path / "vae.json").write_text(json.dumps(L.get_diffusers_vae_config()))
    names = ["red", "green", "blue", "dark"]
    (tmp_path / "tags.csv").write_text("name\n" + "\n".join(names) + "\n")
    g = torch.Generator().manual_seed(3)
    data = {}
    for i in range(12):
        c = i % 3
        arr = torch.randint(0, 60, (72, 80, 3), generator=g, dtype=torch.uint8)
        arr[..., c] += 150
        path = tmp_path / f"im{i}.png"
        Image.fromarray(arr.numpy()).save(path)
        data[str(path)] = f"{names[c]}:1.0, dark:0.5" if i % 4 == 0 else names[c]
    (tmp_path / "data.json").write_text(json.dumps(data))
    out_dir = tmp_path / "out"
    hist = train_full.main(["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path", str(tmp_path / "vae.json"),
                            "--json_path", str(tmp_path / "data.json"), "--tags_csv_path", str(tmp_path / "tags.csv"),
                            "--output_dir", str(out_dir), "--resolution", "64", "--train_batch_size", "2", "--num_epochs", "3",
                            "--num_workers", "0", "--lr_warmup_steps", "1", "--learning_rate", "2e-4", "--logging_steps", "2",
                            "--save_steps", "1"] + extra)
    vals = torch.tensor(hist["train_loss"] + hist["val_loss"])
    assert len(hist["train_loss"]) == 3 and torch.isfinite(vals).all()
    assert json.loads((out_dir / "training_history.json").read_text()) == hist
    tuned = load_file(str(out_dir / "vae" / "diffusion_pytorch_model.safetensors"))
    ref = oracle.state_dict()
    moved = [k for k in ref if k.startswith("encoder.") and not torch.equal(tuned[k], ref[k])]
    assert len(moved) == 106, len(moved)       # every encoder tensor received a gradient and an AdamW update
    assert (out_dir / "vae" / "config.json").exists() and (out_dir / "decoder" / "pytorch_model.bin").exists()
    back = L.load_diffusers_vae_from_pretrained(str(out_dir / "vae"))
    assert back is not None and torch.equal(back.state_dict()["encoder.conv_in.weight"], tuned["encoder.conv_in.weight"])


@pytest.mark.parametrize("extra", [[], ["--use_kl_loss", "--similarity_type", "euclidean"], ["--mixed_precision", "no"]])
def test_train_vae_cli_end_to_end(tmp_path, extra):
    """train_vae.py's command line (reference step train_vae.py:124-186): reconstruction + (KL) + triplet, back-propagated
    natively through the VAE decoder, the posterior samples and the three encoder forwards.  Two epochs on a tiny
    dataset: losses finite, the reconstruction error falls, every encoder AND decoder tensor has moved, and the saved
    VAE loads back."""
    from PIL import Image
    from safetensors.torch import load_file, save_file

    from oracle.decoder import make_oracle_decoder
    from vae_tagger_b200 import train_vae

    oracle, odec = make_oracle_vae(0), make_oracle_decoder(0)
    sd = {k: v.contiguous() for k, v in oracle.state_dict().items()}
    sd.update({"decoder." + k: v.contiguous() for k, v in odec.state_dict().items()})
    save_file(sd, str(tmp_path / "vae.safetensors"))
    (tmp_path / "vae.json").write_text(json.dumps(L.get_diffusers_vae_config()))
    names = ["red", "green", "blue"]
    (tmp_path / "tags.csv").write_text("name\n" + "\n".join(names) + "\n")
    g = torch.Generator().manual_seed(3)
    data = {}
    for i in range(12):
        c = i % 3
        arr = torch.randint(0, 60, (64, 64, 3), generator=g, dtype=torch.uint8)
        arr[..., c] += 150
        path = tmp_path / f"im{i}.png"
        Image.fromarray(arr.numpy()).save(path)
        data[str(path)] = names[c]
    (tmp_path / "data.json").write_text(json.dumps(data))
    out_dir = tmp_path / "out"
    hist = train_vae.main(["--vae_checkpoint", str(tmp_path / "vae.safetensors"), "--vae_config_path", str(tmp_path / "vae.json"),
                           "--json_path", str(tmp_path / "data.json"), "--tags_csv_path", str(tmp_path / "tags.csv"),
                           "--output_dir", str(out_dir), "--resolution", "64", "--train_batch_size", "2", "--num_epochs", "3",
                           "--num_workers", "0", "--lr_warmup_steps", "1", "--learning_rate", "3e-4", "--logging_steps", "2",
                           "--save_steps", "1", "--reconstruction_weight", "1.0"] + extra)
    vals = torch.tensor(hist["train_loss"] + hist["val_loss"])
    assert len(hist["train_loss"]) == 3 and torch.isfinite(vals).all()
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    tuned = load_file(str(out_dir / "vae_checkpoint_epoch_2" / "diffusion_pytorch_model.safetensors"))
    moved = [k for k in sd if not torch.equal(tuned[k], sd[k])]
    assert len(moved) == 106 + 138, len(moved)
    assert L.load_diffusers_vae_from_pretrained(str(out_dir / "best_vae")) is not None
