"""CPU: pin the oracle.  The head / loss / wrapper restatements are checked against golden vectors
produced by the reference's OWN code (tests/golden/make_golden.py, run in the build container where
/root/reference exists).  The encoder restatement is "parity unpinned" by the reference (diffusers is
not installable here) and is pinned by known answers only."""
import math

import pytest
import torch

from oracle import encoder as OE
from oracle import head as OH


def full_sd(golden, case):
    sd = dict(golden["attention_head_base"])
    sd.update(golden["attention_head"][case]["state_dict"])
    return sd


@pytest.mark.parametrize("case", ["att_T11_64x64", "att_T37_40x24", "att_T1000_16x16"])
def test_head_oracle_matches_reference_outputs(golden, case):
    c = golden["attention_head"][case]
    sd = full_sd(golden, case)
    sa = OH.spatial_attention(sd, c["latent"])
    assert torch.allclose(sa, c["spatial"], atol=1e-6)
    fc = OH.feature_compress(sd, sa)
    assert torch.allclose(fc, c["compressed"], atol=1e-6)
    at = OH.self_attention(sd, fc)
    assert torch.allclose(at, c["attended"], atol=1e-6)
    logits = OH.attention_decoder_logits(sd, c["latent"])
    assert torch.allclose(logits, c["logits"], atol=1e-5)
    conf, idx = OH.get_confidence(logits)
    assert torch.allclose(conf, c["conf"], atol=1e-6)
    assert (idx == c["idx"]).float().mean() > 0.99


def test_plain_head_oracle(golden):
    c = golden["plain_head"]
    assert torch.allclose(OH.plain_decoder_logits(c["state_dict"], c["latent"]), c["logits"], atol=1e-5)


def test_focal_oracle(golden):
    f = golden["focal"]
    for (a, g), want in f["cases"].items():
        assert torch.allclose(OH.focal_loss(f["logits"], f["targets"], a, g), want["loss"], atol=1e-7)
        assert torch.allclose(OH.focal_loss_grad(f["logits"], f["targets"], a, g), want["grad"], atol=1e-7)
    # known answer: FocalLoss(1,2) at logit 0 = 0.25*ln2
    assert abs(f["at_zero"].item() - 0.25 * math.log(2)) < 1e-7
    assert abs(OH.focal_loss(torch.zeros(4, 7), torch.ones(4, 7)).item() - 0.173286795) < 1e-7


def test_wrapper_scale_shift(golden):
    w = golden["wrapper"]
    assert torch.allclose(w["mean"] * 0.3611 + 0.1159, w["latent"], atol=1e-7)


def test_head_param_count_and_keys(golden):
    assert golden["att_T1000_param_count"] == 1_443_666 == 1_186_666 + 257 * 1000


def test_encoder_known_answers():
    vae = OE.make_oracle_vae(seed=0)
    sd = vae.state_dict()
    assert sum(p.numel() for p in vae.parameters()) == 34_274_208
    assert len(sd) == 106
    assert sd["encoder.conv_in.weight"].shape == (128, 3, 3, 3)
    assert sd["encoder.conv_out.weight"].shape == (32, 512, 3, 3)
    assert sd["encoder.down_blocks.1.resnets.0.conv_shortcut.weight"].shape == (256, 128, 1, 1)
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in sd
    x = OE.synthetic_images(1, 64, 64)
    with torch.no_grad():
        moments = vae.encoder(x)
        lat = OE.oracle_wrapper_encode(vae, x)
    assert moments.shape == (1, 32, 8, 8) and lat.shape == (1, 16, 8, 8)
    d = OE.OracleDiagonalGaussian(moments)
    assert torch.equal(d.mode(), moments[:, :16])
    assert torch.allclose(lat, moments[:, :16] * 0.3611 + 0.1159)
    noise = torch.randn(1, 16, 8, 8, generator=torch.Generator().manual_seed(0))
    assert torch.allclose(d.sample(noise=noise), d.mean + torch.exp(0.5 * d.logvar) * noise)
    assert d.kl().shape == (1,)
    # determinism of the generators
    assert torch.equal(OE.synthetic_images(2, 8, 8)[1], OE.synthetic_images(3, 8, 8)[1])
    assert torch.equal(OE.make_oracle_vae(0).state_dict()["encoder.conv_in.bias"], sd["encoder.conv_in.bias"])


def test_downsample_pads_right_and_bottom_only():
    d = OE.OracleDownsample2D(4)
    with torch.no_grad():
        d.conv.weight.zero_(); d.conv.bias.zero_()
        d.conv.weight[:, :, 2, 2] = 1.0  # only the bottom-right tap
        x = torch.ones(1, 4, 4, 4)
        y = d(x)
    # last output row/column read the zero padding
    assert y.shape == (1, 4, 2, 2)
    assert y[0, 0, 0, 0] == 4 and y[0, 0, 1, 1] == 0 and y[0, 0, 0, 1] == 0


# ----------------------------------------------------------------------------- training step
def check_digest(got: torch.Tensor, want: dict, atol, rtol=1e-4):
    if "full" in want:
        torch.testing.assert_close(got, want["full"], atol=atol, rtol=rtol)
    else:
        torch.testing.assert_close(got[:8], want["rows"], atol=atol, rtol=rtol)
        torch.testing.assert_close(got.sum(1), want["rowsum"], atol=atol * 30, rtol=rtol)
        torch.testing.assert_close(got.sum(0), want["colsum"], atol=atol * 30, rtol=rtol)


@pytest.mark.parametrize("case", ["nodrop_3x20x36", "drop_4x32x32", "bce_2x16x24"])
def test_head_train_oracle_matches_reference_autograd(golden, train_golden, case):
    c = train_golden["attention"][case]
    sd = full_sd(golden, "att_T11_64x64")
    masks = c["masks"]
    r = OH.head_train_step(sd, c["latent"], c["targets"], c["alpha"], c["gamma"],
                           attn_mask=None if masks is None else masks[0],
                           cls_masks=None if masks is None else masks[1:])
    torch.testing.assert_close(r["logits"], c["logits"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(r["loss"], c["loss"], atol=1e-7, rtol=1e-5)
    torch.testing.assert_close(r["running_mean"], c["running_mean"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(r["running_var"], c["running_var"], atol=1e-6, rtol=1e-5)
    assert list(r["grads"]) == c["param_order"]
    for k, want in c["grads"].items():
        check_digest(r["grads"][k], want, atol=1e-7)


def test_plain_head_train_oracle(train_golden, golden):
    c = train_golden["plain"]
    r = OH.head_train_step(golden["plain_head"]["state_dict"], c["latent"], c["targets"], kind="plain")
    torch.testing.assert_close(r["logits"], c["logits"], atol=1e-5, rtol=1e-5)
    assert list(r["grads"]) == c["param_order"]
    for k, want in c["grads"].items():
        check_digest(r["grads"][k], want, atol=1e-7)


def test_adamw_oracle(train_golden):
    a = train_golden["adamw"]
    p, m, v = a["p0"], torch.zeros(1000), torch.zeros(1000)
    for i, g in enumerate(a["grads"]):
        p, m, v, norm = OH.adamw_step(p, g, m, v, a["lr"], wd=a["wd"], step=i + 1, max_norm=a["max_norm"])
        torch.testing.assert_close(norm, a["norms"][i], atol=1e-6, rtol=1e-6)
        torch.testing.assert_close(p, a["params"][i], atol=1e-7, rtol=1e-6)


# ----------------------------------------------------------------------------- preprocessing
def test_resample_matches_pillow():
    """oracle/resample.py restates Pillow's 8-bit ImagingResample (the arithmetic behind SmartResize's
    ``img.resize(..., LANCZOS)``, modules.py:173, and ``transforms.Resize`` BILINEAR, modules.py:135):
    bit-exact against Pillow itself on seeded images -- down- and up-scaling, one-axis-only, crops."""
    import numpy as np
    from PIL import Image

    from oracle import resample as R

    rng = np.random.default_rng(0)
    cases = [((97, 131), (64, 64)), ((300, 200), (128, 192)), ((64, 48), (128, 96)), ((200, 100), (200, 64)),
             ((100, 200), (64, 200)), ((513, 767), (576, 832)), ((33, 47), (33, 47)),
             ((2, 300), (64, 3)), ((2, 300), (3, 150)), ((2, 201), (2, 100)), ((3, 300), (64, 3)), ((1, 1), (5, 7))]
    for (w, h), (tw, th) in cases:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for kind, pil in ((R.LANCZOS, Image.LANCZOS), (R.BILINEAR, Image.BILINEAR)):
            want = np.asarray(Image.fromarray(img).resize((tw, th), pil))
            assert np.array_equal(R.resize_u8(img, tw, th, kind), want), ((w, h), (tw, th), kind)
    # SmartResize: crop to the bucket ratio, then LANCZOS (modules.py:142-178)
    from vae_tagger_b200.modules import SmartResize

    for (w, h), (tw, th) in [((640, 360), (576, 832)), ((300, 500), (768, 512)), ((512, 512), (512, 512)),
                             ((401, 399), (1024, 1024))]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(SmartResize(tw, th)(Image.fromarray(img)))
        assert np.array_equal(R.smart_resize_u8(img, tw, th), want), ((w, h), (tw, th))


def test_class_balanced_loss_oracle_and_module(train_golden):
    """ClassBalancedLoss (improved_losses.py:58-72): the oracle and the package's module (pure PyTorch on CPU)
    against the reference's own value and gradient."""
    from vae_tagger_b200.improved_losses import ClassBalancedCriterion, ClassBalancedLoss

    g = train_golden["class_balanced"]
    for fn in (lambda x: OH.class_balanced_loss(x, g["targets"], g["samples_per_class"]),
               lambda x: ClassBalancedLoss()(x, g["targets"], g["samples_per_class"]),
               lambda x: ClassBalancedCriterion(g["samples_per_class"])(x, g["targets"])):
        x = g["logits"].clone().requires_grad_(True)
        loss = fn(x)
        loss.backward()
        torch.testing.assert_close(loss.detach(), g["loss"], atol=1e-7, rtol=1e-6)
        torch.testing.assert_close(x.grad, g["grad"], atol=1e-8, rtol=1e-6)


def test_cross_attention_head_oracle(golden, train_golden):
    """--use_cross_attention (modules.py:388-395, :450-459) against the reference module's own logits."""
    c = train_golden["cross_attention_head"]
    sd = full_sd(golden, "att_T11_64x64")
    sd.update(c["extra_state_dict"])
    logits = OH.attention_decoder_logits(sd, c["latent"], use_cross_attention=True)
    torch.testing.assert_close(logits, c["logits"], atol=1e-5, rtol=1e-5)
    assert not torch.allclose(OH.attention_decoder_logits(sd, c["latent"]), c["logits"], atol=1e-3)   # the branch matters


def test_cross_attention_head_train_oracle(golden, train_golden):
    """The --use_cross_attention head in train mode: oracle gradients against the reference's own autograd graph."""
    c = train_golden["cross_attention_head"]
    t = c["train"]
    sd = full_sd(golden, "att_T11_64x64")
    sd.update(c["extra_state_dict"])
    r = OH.head_train_step(sd, c["latent"], t["targets"], use_cross_attention=True)
    torch.testing.assert_close(r["logits"], t["logits"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(r["loss"], t["loss"], atol=1e-7, rtol=1e-5)
    assert sorted(r["grads"]) == sorted(t["param_order"])
    for k, want in t["grads"].items():
        check_digest(r["grads"][k], want, atol=1e-7)


def test_plain_head_without_pooling_oracle(golden, train_golden):
    """ClassificationDecoder(use_adaptive_pooling=False) (modules.py:316-317): eval logits and train-mode gradients."""
    c = train_golden["plain_flat"]
    sd = golden["plain_head"]["state_dict"]
    torch.testing.assert_close(OH.plain_decoder_logits(sd, c["latent"], use_adaptive_pooling=False), c["logits_eval"],
                               atol=1e-5, rtol=1e-5)
    r = OH.head_train_step(sd, c["latent"], c["targets"], kind="plain", use_adaptive_pooling=False)
    torch.testing.assert_close(r["logits"], c["logits"], atol=1e-5, rtol=1e-5)
    for k, want in c["grads"].items():
        check_digest(r["grads"][k], want, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------
# Cross-check of the encoder / decoder restatements (diffusers is not installable, so no true pin exists): a
# SECOND restatement, written independently as pure ``torch.nn.functional`` calls that walk the state dict by
# its diffusers key names, with the attention going through
# ``F.scaled_dot_product_attention`` -- the call diffusers' AttnProcessor2_0 makes -- instead of the oracle's
# explicit matmul / softmax.  Both must agree to fp32 round-off.
def _gn(sd, p, x):
    import torch.nn.functional as F
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd, p, x, stride=1, padding=1):
    import torch.nn.functional as F
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _resnet(sd, p, x):
    import torch.nn.functional as F
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h)))
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def _attention(sd, p, x):
    import torch.nn.functional as F
    b, c, h, w = x.shape
    t = _gn(sd, p + ".group_norm", x).flatten(2).transpose(1, 2)
    q, k, v = (F.linear(t, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]).view(b, -1, 1, c).transpose(1, 2)
               for n in ("to_q", "to_k", "to_v"))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(b, -1, c)
    o = F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return o.transpose(-1, -2).reshape(b, c, h, w) + x


def _mid(sd, p, x):
    x = _resnet(sd, p + ".resnets.0", x)
    x = _attention(sd, p + ".attentions.0", x)
    return _resnet(sd, p + ".resnets.1", x)


def functional_encoder_moments(sd, x):
    import torch.nn.functional as F
    h = _conv(sd, "encoder.conv_in", x)
    for i in range(4):
        for j in range(2):
            h = _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if i < 3:
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _mid(sd, "encoder.mid_block", h)
    return _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h)))


def functional_decoder_image(sd, z):
    import torch.nn.functional as F
    h = _conv(sd, "conv_in", z)
    h = _mid(sd, "mid_block", h)
    for i in range(4):
        for j in range(3):
            h = _resnet(sd, f"up_blocks.{i}.resnets.{j}", h)
        if i < 3:
            h = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "conv_out", F.silu(_gn(sd, "conv_norm_out", h)))


@pytest.mark.parametrize("h,w", [(64, 64), (96, 160)])
def test_encoder_oracle_agrees_with_independent_functional_restatement(h, w):
    vae = OE.make_oracle_vae(seed=0)
    sd = vae.state_dict()
    x = torch.cat([OE.synthetic_images(1, h, w), OE.structured_images(1, h, w)])
    with torch.no_grad():
        want = functional_encoder_moments(sd, x)
        got = vae.encoder(x)
        lat = OE.oracle_wrapper_encode(vae, x)
    assert got.shape == want.shape == (2, 32, h // 8, w // 8)
    assert ((got - want).norm() / want.norm()).item() < 1e-6
    # wrapper: mode() * 0.3611 + 0.1159 (diffusers_vae_loader.py:80-84) on the first 16 moment channels
    assert torch.allclose(lat, want[:, :16] * 0.3611 + 0.1159, atol=1e-6)


def test_oracle_attention_is_sdpa():
    """``OracleAttention`` (explicit matmul / softmax) against the SDPA form on its own weights."""
    torch.manual_seed(3)
    att = OE.OracleAttention(64, 32)
    sd = {"a." + k: v for k, v in att.state_dict().items()}
    x = torch.randn(2, 64, 12, 20)
    with torch.no_grad():
        assert torch.allclose(att(x), _attention(sd, "a", x), atol=2e-6, rtol=1e-5)


def test_decoder_oracle_agrees_with_independent_functional_restatement():
    from oracle import decoder as OD
    dec = OD.make_oracle_decoder(seed=0)
    z = torch.randn(1, 16, 8, 12, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = functional_decoder_image(dec.state_dict(), z)
        got = dec(z)
    assert got.shape == want.shape == (1, 3, 64, 96)
    assert ((got - want).norm() / want.norm()).item() < 5e-6      # fp32 round-off through 44 layers (measured 1.8e-6)
