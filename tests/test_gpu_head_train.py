"""Head TRAINING step parity on the GPU: ``vt_head_train_step`` (train-mode forward +
focal loss + analytic backward) and ``vt_adamw_step`` against

  * the golden vectors produced by the reference's own modules in train() mode with the reference's
    own autograd graph (tests/golden/make_train_golden.py), where dropout is off, and
  * the oracle restatement (oracle/head.py ``head_train_step``) with the SAME dropout masks the
    kernels use (``vt_head_dropout_masks``), where dropout is on.

fp32 kernels vs fp32 CPU autograd.  Tolerances: logits 2e-5 relative L2; every gradient tensor 1e-4
relative L2 (different summation orders over up to B*H*W = 131 072 terms) -- except tensors whose
exact value is zero by construction (the conv bias in front of a train-mode BatchNorm; the key
projection's bias, which shifts every score of a softmax row by the same amount), which are
compared against the size of the neighbouring weight gradient.
"""
import pytest
import torch

from oracle import head as OH
from vae_tagger_b200 import _native

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def full_sd(golden, case="att_T11_64x64"):
    sd = dict(golden["attention_head_base"])
    sd.update(golden["attention_head"][case]["state_dict"])
    return sd


def setup_head(ctx, sd, kind=_native.HEAD_ATTENTION, **cfg):
    T = sd["classifier.12.weight" if kind == _native.HEAD_ATTENTION else "classifier.8.weight"].shape[0]
    ctx.configure_head(kind, latent_channels=16, num_classes=T, **cfg)
    layout = ctx.head_param_layout()
    total = layout[-1][1] + layout[-1][2]
    flat = torch.empty(total, device="cuda")
    for name, off, n in layout:
        assert sd[name].numel() == n, name
        flat[off:off + n] = sd[name].reshape(-1).cuda()
    return layout, flat


def unflatten(layout, flat, sd):
    return {name: flat[off:off + n].view(sd[name].shape).cpu() for name, off, n in layout}


ZERO_BY_CONSTRUCTION = {"feature_compress.0.bias": "feature_compress.0.weight",
                        "self_attention_post.k_proj.bias": "self_attention_post.k_proj.weight"}


def compare_grads(got, want, tol=1e-4):
    assert list(got) == list(want)
    bad = {}
    for k in want:
        if k in ZERO_BY_CONSTRUCTION:
            scale = want[ZERO_BY_CONSTRUCTION[k]].abs().max().item()
            if got[k].abs().max().item() > 1e-4 * max(scale, 1e-12):
                bad[k] = ("zero-by-construction", got[k].abs().max().item(), scale)
            continue
        r = rel(got[k], want[k])
        if not r < tol:
            bad[k] = (r, got[k].flatten()[:4].tolist(), want[k].flatten()[:4].tolist())
    assert not bad, bad


def check_digest(got, want, tol=1e-4, key=""):
    """Golden gradients of the big classifier matrices are stored as digests (first 8 rows, row sums,
    column sums).  A column sum can be zero by construction (the gradient w.r.t. a LayerNorm input sums
    to zero over the features), so sums are compared against the size of the summed entries."""
    if "full" in want:
        assert rel(got, want["full"]) < tol, key
        return
    assert rel(got[:8], want["rows"]) < tol, key
    scale = want["rows"].abs().max().item()
    for dim, name in ((1, "rowsum"), (0, "colsum")):
        err = (got.sum(dim) - want[name]).norm().item()
        allowed = tol * (want[name].norm().item() + scale * want[name].numel() ** 0.5)
        assert err <= allowed, (key, name, err, allowed)


def check_golden_grads(got, golden_grads):
    for k, want in golden_grads.items():
        if k in ZERO_BY_CONSTRUCTION:
            assert got[k].abs().max().item() <= 1e-4 * got[ZERO_BY_CONSTRUCTION[k]].abs().max().item(), k
        else:
            check_digest(got[k], want, key=k)


def test_layout_is_the_reference_parameter_order(ctx, golden, train_golden):
    sd = full_sd(golden)
    layout, _ = setup_head(ctx, sd)
    assert [n for n, _, _ in layout] == train_golden["attention"]["nodrop_3x20x36"]["param_order"]
    assert layout[-1][1] + layout[-1][2] == 1_186_666 + 257 * 11
    layout, _ = setup_head(ctx, golden["plain_head"]["state_dict"], kind=_native.HEAD_PLAIN)
    assert [n for n, _, _ in layout] == train_golden["plain"]["param_order"]


@pytest.mark.parametrize("case", ["nodrop_3x20x36", "bce_2x16x24"])
def test_train_step_vs_reference_golden(ctx, golden, train_golden, case):
    """Dropout off: straight against the reference's own train-mode forward and autograd gradients."""
    c = train_golden["attention"][case]
    sd = full_sd(golden)
    layout, flat = setup_head(ctx, sd)
    grads = torch.zeros_like(flat)
    rm, rv = sd["feature_compress.1.running_mean"].cuda(), sd["feature_compress.1.running_var"].cuda()
    nbt = sd["feature_compress.1.num_batches_tracked"].cuda()
    nbt0 = nbt.item()
    loss, logits = ctx.head_train_step(c["latent"].cuda(), c["targets"].cuda(), flat, grads, rm, rv, nbt,
                                       focal_alpha=c["alpha"], focal_gamma=c["gamma"], dropout=False,
                                       want_logits=True)
    assert rel(logits.cpu(), c["logits"]) < 2e-5
    assert abs(loss.item() - c["loss"].item()) < 1e-6 * max(1.0, abs(c["loss"].item()))
    assert rel(rm.cpu(), c["running_mean"]) < 1e-5 and rel(rv.cpu(), c["running_var"]) < 1e-5
    assert nbt.item() == nbt0 + 1 == c["num_batches_tracked"].item()
    got = unflatten(layout, grads, sd)
    check_golden_grads(got, c["grads"])


@pytest.mark.parametrize("B,lh,lw,T_case,seed", [(4, 32, 32, "att_T11_64x64", 5), (3, 20, 36, "att_T37_40x24", 6),
                                                 (8, 128, 128, "att_T1000_16x16", 7)])
def test_train_step_with_dropout_vs_oracle(ctx, golden, B, lh, lw, T_case, seed):
    """Dropout on: the oracle applies the masks of the kernels' counter-based generator."""
    sd = full_sd(golden, T_case)
    T = sd["classifier.12.weight"].shape[0]
    layout, flat = setup_head(ctx, sd)
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(B, 16, lh, lw, generator=g) * 0.5 + 0.1
    tgt = (torch.rand(B, T, generator=g) < 0.1).float()
    attn_mask, cls_masks = ctx.head_dropout_masks(B, 0.1, seed)
    keep = [attn_mask.mean().item()] + [m.mean().item() for m in cls_masks]
    for k, p in zip(keep, (0.1, 0.3, 0.2, 0.1)):
        assert abs(k - (1 - p)) < 0.03, keep      # the generator drops the stated fraction
    grads = torch.zeros_like(flat)
    rm, rv = sd["feature_compress.1.running_mean"].cuda(), sd["feature_compress.1.running_var"].cuda()
    loss, logits = ctx.head_train_step(lat.cuda(), tgt.cuda(), flat, grads, rm, rv, None, dropout=True,
                                       attention_dropout=0.1, seed=seed, want_logits=True)
    # ground truth in double precision: several gradients are cancelling sums over B*H*W terms, where the
    # fp32 CPU autograd graph is itself only good to ~1e-4
    want = OH.head_train_step(sd, lat, tgt, attn_mask=attn_mask.cpu(), cls_masks=[m.cpu() for m in cls_masks],
                              dtype=torch.float64)
    want = {k: ({n: g.float() for n, g in v.items()} if isinstance(v, dict) else v.float()) for k, v in want.items()}
    assert rel(logits.cpu(), want["logits"]) < 2e-5
    assert abs(loss.item() - want["loss"].item()) < 1e-6 * max(1.0, want["loss"].item())
    assert rel(rm.cpu(), want["running_mean"]) < 1e-5 and rel(rv.cpu(), want["running_var"]) < 1e-5
    compare_grads(unflatten(layout, grads, sd), want["grads"])
    # a different seed gives different masks; dropout=False ignores the seed
    other, _ = ctx.head_dropout_masks(B, 0.1, seed + 1)
    assert not torch.equal(other, attn_mask)


@pytest.mark.parametrize("sa,mh", [(False, True), (True, False), (False, False)])
def test_train_step_variants(ctx, golden, sa, mh):
    sd = full_sd(golden)
    layout, flat = setup_head(ctx, sd, use_spatial_attention=sa, use_self_attention=mh)
    names = [n for n, _, _ in layout]
    assert any(n.startswith("spatial_attention.") for n in names) == sa
    assert any(n.startswith("self_attention_post.") for n in names) == mh
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(3, 16, 24, 40, generator=g)
    tgt = (torch.rand(3, 11, generator=g) < 0.3).float()
    grads = torch.zeros_like(flat)
    ctx.head_train_step(lat.cuda(), tgt.cuda(), flat, grads, dropout=False)
    want = OH.head_train_step({k: v for k, v in sd.items() if k in names or "running" in k}, lat, tgt,
                              use_spatial_attention=sa, use_self_attention=mh)
    compare_grads(unflatten(layout, grads, sd), want["grads"])


def test_plain_head_train_step(ctx, golden, train_golden):
    c = train_golden["plain"]
    sd = golden["plain_head"]["state_dict"]
    layout, flat = setup_head(ctx, sd, kind=_native.HEAD_PLAIN)
    grads = torch.zeros_like(flat)
    loss, logits = ctx.head_train_step(c["latent"].cuda(), c["targets"].cuda(), flat, grads, dropout=False,
                                       want_logits=True)
    assert rel(logits.cpu(), c["logits"]) < 2e-5
    assert abs(loss.item() - c["loss"].item()) < 1e-6
    check_golden_grads(unflatten(layout, grads, sd), c["grads"])


def test_accumulation_scale_and_determinism(ctx, golden):
    sd = full_sd(golden)
    layout, flat = setup_head(ctx, sd)
    g = torch.Generator().manual_seed(9)
    lat = (torch.randn(4, 16, 32, 32, generator=g)).cuda()
    tgt = (torch.rand(4, 11, generator=g) < 0.3).float().cuda()
    runs = []
    for _ in range(2):
        grads = torch.zeros_like(flat)
        loss, _ = ctx.head_train_step(lat, tgt, flat, grads, dropout=True, seed=42)
        runs.append((grads.clone(), loss.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])   # bit-reproducible
    # two accumulated half-scale steps == one full step (gradient accumulation, train_decoder.py:190)
    acc = torch.zeros_like(flat)
    loss = torch.zeros(1, device="cuda")
    for _ in range(2):
        ctx.head_train_step(lat, tgt, flat, acc, dropout=True, seed=42, loss_scale=0.5, loss=loss)
    assert rel(acc, runs[0][0]) < 1e-6
    assert abs(loss.item() - runs[0][1].item()) < 1e-6
    # forward + loss only
    l2, _ = ctx.head_train_step(lat, tgt, flat, None, dropout=True, seed=42)
    assert torch.equal(l2, runs[0][1])


def test_adamw_step_vs_torch_golden(ctx, train_golden):
    a = train_golden["adamw"]
    p = a["p0"].cuda().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    norm = torch.zeros(1, device="cuda")
    for i, g in enumerate(a["grads"]):
        gd = g.cuda().clone()
        ctx.adamw_step(p, gd, m, v, lr=a["lr"], weight_decay=a["wd"], step=i + 1, max_norm=a["max_norm"],
                       norm_out=norm)
        assert abs(norm.item() - a["norms"][i].item()) < 1e-5 * a["norms"][i].item()
        assert (p.cpu() - a["params"][i]).abs().max().item() < 2e-7
        assert gd.abs().max().item() == 0.0          # zero_grad fused
    # grad_scale (1/world after the all-reduce) and no clipping
    p2 = a["p0"].cuda().clone()
    m2, v2 = torch.zeros_like(p2), torch.zeros_like(p2)
    ctx.adamw_step(p2, (a["grads"][1] * 4).cuda(), m2, v2, lr=1e-3, weight_decay=0.0, step=1, grad_scale=0.25,
                   max_norm=0.0)
    want, _, _, _ = OH.adamw_step(a["p0"], a["grads"][1], torch.zeros(1000), torch.zeros(1000), 1e-3, wd=0.0, step=1)
    assert (p2.cpu() - want).abs().max().item() < 2e-7


def test_decoder_trainer_native_step_matches_autograd_step(golden):
    """DecoderTrainer (train_decoder.py:178-206 drop-in): the native step and the PyTorch-autograd step
    of the same module give the same parameters after three optimizer steps (dropout off)."""
    from vae_tagger_b200 import modules as M
    from vae_tagger_b200.improved_losses import FocalLoss
    from vae_tagger_b200.train_decoder import DecoderTrainer

    class FrozenLatent(torch.nn.Module):          # stands in for the frozen encoder: encode() = identity
        def encode(self, x):
            return x

    sd = full_sd(golden)
    g = torch.Generator().manual_seed(21)
    lat = [torch.randn(4, 16, 32, 32, generator=g).cuda() for _ in range(3)]
    tgt = [(torch.rand(4, 11, generator=g) < 0.3).float().cuda() for _ in range(3)]
    finals = []
    for native in (True, False):
        dec = M.create_attention_decoder(16, 32, 32, 11, attention_config={"attention_dropout": 0.0})
        dec.load_state_dict(sd)
        dec = dec.cuda()
        for mod in dec.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        opt = torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-2)
        tr = DecoderTrainer(FrozenLatent(), dec, FocalLoss(1.0, 2.0), opt, None, max_grad_norm=1.0,
                            native_step=native)
        assert tr.native == native
        losses = [tr.step(x, y).item() for x, y in zip(lat, tgt)]
        tr.flush()
        finals.append(({k: v.detach().cpu().clone() for k, v in dec.state_dict().items()}, losses))
    (sd_n, l_n), (sd_a, l_a) = finals
    assert abs(l_n[0] - l_a[0]) < 1e-6, (l_n, l_a)          # same parameters: same loss
    assert max(abs(a - b) for a, b in zip(l_n, l_a)) < 1e-4, (l_n, l_a)
    bad = {}
    for k in sd_a:
        if k.endswith("num_batches_tracked"):
            assert sd_n[k].item() == sd_a[k].item() == 3 + sd[k].item()
        elif k in ZERO_BY_CONSTRUCTION:           # zero gradient by construction: Adam amplifies the noise
            continue
        else:
            # Adam normalises every gradient entry: an entry that is rounding noise in both paths moves by
            # +-lr per step in either, so compare the mean movement, and bound the worst entry by 2*steps*lr
            d = (sd_n[k] - sd_a[k]).abs()
            if d.mean().item() > 1e-4 or d.max().item() > 6.5e-3:
                bad[k] = (d.mean().item(), d.max().item())
    assert not bad, bad


def test_edge_shapes_and_errors(ctx, golden):
    """Smallest batch / latent the train-mode BatchNorm allows, a ragged non-multiple-of-8 latent, and the
    argument errors of the C-ABI."""
    sd = full_sd(golden)
    layout, flat = setup_head(ctx, sd)
    g = torch.Generator().manual_seed(4)
    for B, lh, lw in ((1, 8, 8), (2, 9, 13), (33, 8, 8)):
        lat = torch.randn(B, 16, lh, lw, generator=g)
        tgt = (torch.rand(B, 11, generator=g) < 0.3).float()
        grads = torch.zeros_like(flat)
        loss, logits = ctx.head_train_step(lat.cuda(), tgt.cuda(), flat, grads, dropout=False, want_logits=True)
        want = OH.head_train_step(sd, lat, tgt, dtype=torch.float64)
        assert rel(logits.cpu().double(), want["logits"]) < 2e-5
        compare_grads(unflatten(layout, grads, sd),
                      {k: v.float() for k, v in want["grads"].items()}, tol=2e-4)
    with pytest.raises(_native.NativeError):
        ctx.head_train_step(torch.zeros(1, 16, 1, 1).cuda(), torch.zeros(1, 11).cuda(), flat, torch.zeros_like(flat))
    with pytest.raises(_native.NativeError):
        ctx.head_train_step(torch.zeros(2, 16, 8, 8).cuda(), torch.zeros(2, 11).cuda(), flat, None,
                            attention_dropout=1.0)


def test_class_balanced_training_step(ctx, golden, train_golden):
    """--use_class_balanced (train_decoder.py:188-189): the class weights go into the loss kernel; gradients
    against the oracle's ClassBalancedLoss step, and DecoderTrainer picks the native step for it."""
    from vae_tagger_b200.improved_losses import ClassBalancedCriterion, class_balanced_weights

    sd = full_sd(golden)
    layout, flat = setup_head(ctx, sd)
    spc = train_golden["class_balanced"]["samples_per_class"]
    g = torch.Generator().manual_seed(12)
    lat = torch.randn(3, 16, 16, 24, generator=g)
    tgt = (torch.rand(3, 11, generator=g) < 0.3).float()
    w = torch.tensor(class_balanced_weights(spc), dtype=torch.float32).cuda()
    grads = torch.zeros_like(flat)
    loss, logits = ctx.head_train_step(lat.cuda(), tgt.cuda(), flat, grads, focal_alpha=1.0, focal_gamma=0.0,
                                       dropout=False, class_weights=w, want_logits=True)
    want = OH.head_train_step(sd, lat, tgt, samples_per_class=spc)
    assert abs(loss.item() - want["loss"].item()) < 1e-6 * max(1.0, want["loss"].item())
    compare_grads(unflatten(layout, grads, sd), want["grads"])

    from vae_tagger_b200 import modules as M
    from vae_tagger_b200.train_decoder import DecoderTrainer

    class FrozenLatent(torch.nn.Module):
        def encode(self, x):
            return x

    dec = M.create_attention_decoder(16, 16, 24, 11, attention_config={})
    dec.load_state_dict(sd)
    dec = dec.cuda()
    opt = torch.optim.AdamW(dec.parameters(), lr=1e-3)
    tr = DecoderTrainer(FrozenLatent(), dec, ClassBalancedCriterion(spc), opt, None, native_step=True)
    assert tr.native and tr._gamma == 0.0 and tr._class_w is not None
    l0 = tr.step(lat.cuda(), tgt.cuda()).item()
    tr.flush()
    assert l0 > 0 and torch.isfinite(torch.tensor(l0))


def test_gradient_accumulation_follows_the_reference_loop(golden):
    """gradient_accumulation_steps = 2: the reference (train_decoder.py:186-203) scales the loss by 1/2, clips the
    accumulated gradient after EVERY backward and steps on every second one.  DecoderTrainer (native step) against
    a literal transcription of that loop on the PyTorch graph of the same module."""
    from vae_tagger_b200 import modules as M
    from vae_tagger_b200.improved_losses import FocalLoss
    from vae_tagger_b200.train_decoder import DecoderTrainer

    class FrozenLatent(torch.nn.Module):
        def encode(self, x):
            return x

    sd = full_sd(golden)
    g = torch.Generator().manual_seed(31)
    lat = [torch.randn(2, 16, 16, 16, generator=g).cuda() * 3 for _ in range(4)]     # large: the clip is active
    tgt = [(torch.rand(2, 11, generator=g) < 0.3).float().cuda() for _ in range(4)]

    def fresh():
        dec = M.create_attention_decoder(16, 16, 16, 11, attention_config={"attention_dropout": 0.0})
        dec.load_state_dict(sd)
        dec = dec.cuda()
        for mod in dec.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return dec, torch.optim.AdamW(dec.parameters(), lr=1e-3, weight_decay=1e-2)

    max_norm = 0.05
    dec_n, opt_n = fresh()
    tr = DecoderTrainer(FrozenLatent(), dec_n, FocalLoss(1.0, 2.0), opt_n, None, max_grad_norm=max_norm,
                        gradient_accumulation_steps=2, native_step=True)
    for x, y in zip(lat, tgt):
        tr.step(x, y)
    tr.flush()

    dec_r, opt_r = fresh()
    loss_fn = FocalLoss(1.0, 2.0)
    dec_r.train()
    for step, (x, y) in enumerate(zip(lat, tgt)):
        with torch.enable_grad():
            loss = loss_fn(dec_r(x), y) / 2
        loss.backward()
        torch.nn.utils.clip_grad_norm_(dec_r.parameters(), max_norm)
        if (step + 1) % 2 == 0:
            opt_r.step()
            opt_r.zero_grad()
    bad = {}
    for (k, a), (_, b) in zip(dec_n.state_dict().items(), dec_r.state_dict().items()):
        if k in ZERO_BY_CONSTRUCTION or k.endswith("num_batches_tracked"):
            continue
        d = (a - b).abs()
        if d.mean().item() > 1e-4 or d.max().item() > 4.5e-3:
            bad[k] = (d.mean().item(), d.max().item())
    assert not bad, bad


def test_cross_attention_training_step(ctx, golden, train_golden):
    """--use_cross_attention in training: logits, loss and every gradient (incl. cross_attention.* and
    query_generator.*) against the reference module's own train-mode forward and autograd gradients."""
    c = train_golden["cross_attention_head"]
    t = c["train"]
    sd = full_sd(golden)
    sd.update(c["extra_state_dict"])
    layout, flat = setup_head(ctx, sd, use_cross_attention=True)
    assert [n for n, _, _ in layout] == t["param_order"]
    grads = torch.zeros_like(flat)
    loss, logits = ctx.head_train_step(c["latent"].cuda(), t["targets"].cuda(), flat, grads, dropout=False,
                                       want_logits=True)
    assert rel(logits.cpu(), t["logits"]) < 2e-5
    assert abs(loss.item() - t["loss"].item()) < 1e-6
    got = unflatten(layout, grads, sd)
    zero = dict(ZERO_BY_CONSTRUCTION)
    zero["cross_attention.k_proj.bias"] = "cross_attention.k_proj.weight"     # softmax shift invariance again
    for k, want in t["grads"].items():
        if k in zero:
            assert got[k].abs().max().item() <= 1e-4 * got[zero[k]].abs().max().item(), k
        else:
            check_digest(got[k], want, key=k)
    # and against the fp64 oracle with dropout on
    attn_mask, cls_masks = ctx.head_dropout_masks(3, 0.1, 77)
    grads.zero_()
    ctx.head_train_step(c["latent"].cuda(), t["targets"].cuda(), flat, grads, dropout=True, seed=77)
    want = OH.head_train_step(sd, c["latent"], t["targets"], use_cross_attention=True, attn_mask=attn_mask.cpu(),
                              cls_masks=[m.cpu() for m in cls_masks], dtype=torch.float64)
    wg = {k: v.float() for k, v in want["grads"].items()}
    got = unflatten(layout, grads, sd)
    bad = {}
    for k in wg:
        if k in zero:
            continue
        r = rel(got[k], wg[k])
        if not r < 1e-4:
            bad[k] = r
    assert not bad, bad


def test_plain_head_without_pooling(ctx, golden, train_golden):
    """ClassificationDecoder(use_adaptive_pooling=False): the flattened latent feeds the MLP directly -- module
    inference and the native training step against the reference's own outputs / gradients."""
    from vae_tagger_b200 import modules as M

    c = train_golden["plain_flat"]
    sd = golden["plain_head"]["state_dict"]
    dec = M.ClassificationDecoder(16, 4, 4, 11, use_adaptive_pooling=False)
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    assert rel(dec(c["latent"].cuda()).cpu(), c["logits_eval"]) < 2e-5
    with pytest.raises(_native.NativeError):
        dec(torch.zeros(1, 16, 8, 8).cuda())                 # built for a 4x4 latent
    layout, flat = setup_head(ctx, sd, kind=_native.HEAD_PLAIN, plain_flat_dim=256)
    grads = torch.zeros_like(flat)
    loss, logits = ctx.head_train_step(c["latent"].cuda(), c["targets"].cuda(), flat, grads, dropout=False,
                                       want_logits=True)
    assert rel(logits.cpu(), c["logits"]) < 2e-5 and abs(loss.item() - c["loss"].item()) < 1e-6
    check_golden_grads(unflatten(layout, grads, sd), c["grads"])
