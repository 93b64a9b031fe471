"""Tag-head parity on the GPU against the golden vectors produced by the reference's own
modules.py (tests/golden/make_golden.py) and against the oracle restatement.

fp32 kernels vs fp32 CPU reference: tolerance 2e-5 relative on logits (different summation order
and expf implementations), 1e-6 absolute on sigmoid outputs; sort indices must be identical where
neighbouring confidences differ by more than 1e-6.
"""
import pytest
import torch

from oracle import head as OH
from vae_tagger_b200 import modules as M

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def full_sd(golden, case):
    sd = dict(golden["attention_head_base"])
    sd.update(golden["attention_head"][case]["state_dict"])
    return sd


@pytest.mark.parametrize("case", ["att_T11_64x64", "att_T37_40x24", "att_T1000_16x16"])
def test_attention_head_vs_reference_golden(golden, case):
    c = golden["attention_head"][case]
    sd = full_sd(golden, case)
    T = sd["classifier.12.weight"].shape[0]
    lat = c["latent"]
    dec = M.create_attention_decoder(16, lat.shape[2], lat.shape[3], T, attention_config={})
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    logits = dec(lat.cuda()).cpu()
    assert rel(logits, c["logits"]) < 2e-5, rel(logits, c["logits"])
    conf, idx = dec.get_confidence(lat.cuda())
    conf, idx = conf.cpu(), idx.cpu()
    assert idx.dtype == torch.int64
    assert (conf - c["conf"]).abs().max().item() < 1e-6
    # sorted, a permutation, and consistent with the unsorted probabilities
    assert (conf[:, :-1] >= conf[:, 1:]).all()
    assert torch.equal(idx.sort(dim=1).values, torch.arange(T).expand_as(idx))
    probs = torch.sigmoid(logits)
    assert (torch.gather(probs, 1, idx) - conf).abs().max().item() < 1e-6
    gap = (c["conf"][:, :-1] - c["conf"][:, 1:]).abs()
    safe = torch.ones_like(c["idx"], dtype=torch.bool)
    safe[:, :-1] &= gap > 2e-6
    safe[:, 1:] &= gap > 2e-6
    assert torch.equal(idx[safe], c["idx"][safe])


def test_plain_head_vs_reference_golden(golden):
    c = golden["plain_head"]
    dec = M.create_attention_decoder(16, 64, 64, 11, attention_config=None)
    dec.load_state_dict(c["state_dict"])
    dec = dec.cuda().eval()
    logits = dec(c["latent"].cuda()).cpu()
    assert rel(logits, c["logits"]) < 2e-5


def test_threshold_counts_and_oracle(golden):
    case = "att_T11_64x64"
    c = golden["attention_head"][case]
    sd = full_sd(golden, case)
    dec = M.create_attention_decoder(16, 64, 64, 11, attention_config={})
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    out = dec.tag(c["latent"].cuda(), threshold=0.5)
    conf, idx, cnt = out["conf"].cpu(), out["idx"].cpu(), out["count"].cpu()
    for b in range(conf.shape[0]):
        want = OH.threshold_tags(c["conf"][b], c["idx"][b], 0.5)
        assert int(cnt[b]) == want["total_tags_above_threshold"]
        assert [int(i) for i in idx[b, : int(cnt[b])]] == [i for i, _ in want["predicted"]]


def test_variants_no_spatial_no_self(golden):
    sd = full_sd(golden, "att_T11_64x64")
    lat = golden["attention_head"]["att_T11_64x64"]["latent"]
    for sa, se in ((False, True), (True, False), (False, False)):
        dec = M.create_attention_decoder(16, 64, 64, 11, attention_config={
            "use_spatial_attention": sa, "use_self_attention": se})
        dec.load_state_dict(sd, strict=False)
        dec = dec.cuda().eval()
        got = dec(lat.cuda()).cpu()
        ref = OH.attention_decoder_logits(sd, lat, use_spatial_attention=sa, use_self_attention=se)
        assert rel(got, ref) < 2e-5, (sa, se, rel(got, ref))


def test_focal_loss_native(ctx, golden):
    f = golden["focal"]
    for (a, g), want in f["cases"].items():
        loss, grad = ctx.focal_loss(f["logits"].cuda(), f["targets"].cuda(), alpha=a, gamma=g)
        n = f["logits"].numel()
        assert abs(loss.item() / n - want["loss"].item()) < 1e-6 * max(1.0, abs(want["loss"].item()))
        assert rel(grad.cpu(), want["grad"]) < 1e-5


def test_head_extremes(ctx, golden):
    """Largest supported tag vocabulary (16 384: the sort's shared-memory limit), batch 1 and a batch that is
    not a multiple of the linear kernel's 32-row pass; out-of-range vocabularies are refused."""
    from vae_tagger_b200 import _native

    torch.manual_seed(5)
    for T, B in ((16384, 1), (16384, 3), (1000, 37)):
        dec = M.create_attention_decoder(16, 16, 16, T, attention_config={}).eval()
        sd = {k: v.clone() for k, v in dec.state_dict().items()}
        lat = torch.randn(B, 16, 16, 16)
        want = OH.attention_decoder_logits(sd, lat)
        dec = dec.cuda()
        out = dec.tag(lat.cuda(), threshold=0.5)
        logits = dec(lat.cuda()).cpu()
        assert rel(logits, want) < 2e-5
        conf = out["conf"].cpu()
        assert (conf[:, :-1] >= conf[:, 1:]).all()
        assert torch.equal(out["idx"].cpu().sort(dim=1).values, torch.arange(T).expand(B, T))
        assert torch.equal(out["count"].cpu(), (torch.sigmoid(logits) >= 0.5).sum(1).to(torch.int32))
    with pytest.raises(_native.NativeError):
        ctx.configure_head(_native.HEAD_ATTENTION, latent_channels=16, num_classes=16385)


def test_cross_attention_head_native(golden, train_golden):
    """--use_cross_attention (modules.py:388-395, :450-459) runs as native kernels in inference: against the
    reference module's own logits; it also trains natively."""
    from vae_tagger_b200 import _native
    from vae_tagger_b200.train_decoder import DecoderTrainer

    c = train_golden["cross_attention_head"]
    sd = full_sd(golden, "att_T11_64x64")
    sd.update(c["extra_state_dict"])
    dec = M.create_attention_decoder(16, 24, 40, 11, attention_config={"use_cross_attention": True})
    assert sorted(dec.state_dict().keys()) == sorted(sd.keys())
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    assert dec._use_native()
    logits = dec(c["latent"].cuda()).cpu()
    assert rel(logits, c["logits"]) < 2e-5, rel(logits, c["logits"])
    conf, idx = dec.get_confidence(c["latent"].cuda())
    assert (conf.cpu() - torch.sigmoid(c["logits"]).sort(descending=True).values).abs().max().item() < 1e-6
    # training with the branch: native step (AdamW + BCE), and a non-AdamW optimizer keeps the PyTorch graph
    class FrozenLatent(torch.nn.Module):
        def encode(self, x):
            return x

    y = (torch.rand(3, 11) < 0.3).float().cuda()
    tr = DecoderTrainer(FrozenLatent(), dec, torch.nn.BCEWithLogitsLoss(), torch.optim.AdamW(dec.parameters(), lr=1e-3),
                        None, native_step=True)
    assert tr.native
    losses = [tr.step(c["latent"].cuda(), y).item() for _ in range(6)]
    tr.flush()
    assert losses[-1] < losses[0]
    with pytest.raises(_native.NativeError):
        DecoderTrainer(FrozenLatent(), dec, torch.nn.BCEWithLogitsLoss(), torch.optim.SGD(dec.parameters(), lr=0.1), None,
                       native_step=True)
    tr2 = DecoderTrainer(FrozenLatent(), dec, torch.nn.BCEWithLogitsLoss(), torch.optim.SGD(dec.parameters(), lr=0.1), None)
    assert not tr2.native
    assert tr2.step(c["latent"].cuda(), y).item() > 0
    tr2.flush()
