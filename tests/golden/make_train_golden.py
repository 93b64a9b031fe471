"""Golden vectors of the head TRAINING step from the reference's OWN code (build container only).

    python tests/golden/make_train_golden.py

``modules.py`` / ``improved_losses.py`` are imported unchanged from ``/root/reference`` (stub
``diffusers``, see make_golden.py).  The reference modules run in ``train()`` mode -- batch-statistics
BatchNorm, Dropout active -- and the gradients come from the reference's own autograd graph.
Dropout is made reproducible by replacing ``torch.nn.functional.dropout`` for the duration of the
forward with a function that applies recorded keep-masks in call order (attention dropout first,
then the three classifier dropouts: modules.py:81, :405, :410, :415); the masks are stored.
Also: torch.optim.AdamW + clip_grad_norm_ on a flat tensor (train_decoder.py:197-203).

Parameters are those of ``head_golden.pt`` (case att_T11_64x64 / plain_head), so only inputs, masks and
gradients are stored; gradients of the three big classifier matrices are stored as digests
(first 8 rows, row sums, column sums).

  head_train_golden.pt
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_stub_diffusers  # noqa: E402


def digest(g: torch.Tensor):
    if g.numel() <= 20000:
        return {"full": g.clone()}
    return {"rows": g[:8].clone(), "rowsum": g.sum(1), "colsum": g.sum(0)}


def main():
    install_stub_diffusers()
    sys.path.insert(0, REF)
    import improved_losses as ref_losses  # noqa: E402
    import modules as ref_modules  # noqa: E402
    import torch.nn.functional as F

    base = torch.load(os.path.join(HERE, "head_golden.pt"), map_location="cpu", weights_only=False)
    quiet = contextlib.redirect_stdout(io.StringIO())
    out = {}

    sd = dict(base["attention_head_base"])
    sd.update(base["attention_head"]["att_T11_64x64"]["state_dict"])
    cases = {}
    for name, (B, lh, lw, with_dropout, alpha, gamma, seed) in {
        "nodrop_3x20x36": (3, 20, 36, False, 1.0, 2.0, 10),     # 20/8, 36/8: overlapping pooling windows
        "drop_4x32x32": (4, 32, 32, True, 0.25, 2.0, 11),
        "bce_2x16x24": (2, 16, 24, False, 1.0, 0.0, 12),        # gamma 0 == BCEWithLogitsLoss
    }.items():
        torch.manual_seed(seed)
        with quiet:
            dec = ref_modules.create_attention_decoder(16, lh, lw, 11, attention_config={})
        dec.load_state_dict(sd)
        dec.train()
        lat = torch.randn(B, 16, lh, lw) * 0.5 + 0.1
        tgt = (torch.rand(B, 11) < 0.3).float()
        ps = [0.1, 0.3, 0.2, 0.1]
        shapes = [(B, 8, 64, 64), (B, 1024), (B, 512), (B, 256)]
        masks = [(torch.rand(s) >= p).float() for s, p in zip(shapes, ps)] if with_dropout else None
        calls = []
        real_dropout = F.dropout

        def fake_dropout(x, p=0.5, training=True, inplace=False):
            i = len(calls)
            calls.append((tuple(x.shape), p))
            assert training
            if masks is None:
                return x
            assert tuple(x.shape) == shapes[i] and abs(p - ps[i]) < 1e-9
            return x * masks[i] / (1.0 - p)

        F.dropout = fake_dropout
        try:
            with quiet:
                logits = dec(lat)
        finally:
            F.dropout = real_dropout
        assert [c[1] for c in calls] == ps
        loss = ref_losses.FocalLoss(alpha=alpha, gamma=gamma)(logits, tgt)
        loss.backward()
        cases[name] = {
            "latent": lat, "targets": tgt, "alpha": alpha, "gamma": gamma,
            "masks": masks, "logits": logits.detach().clone(), "loss": loss.detach().clone(),
            "grads": {k: digest(p.grad) for k, p in dec.named_parameters()},
            "running_mean": dec.feature_compress[1].running_mean.clone(),
            "running_var": dec.feature_compress[1].running_var.clone(),
            "num_batches_tracked": dec.feature_compress[1].num_batches_tracked.clone(),
            "param_order": [k for k, _ in dec.named_parameters()],
        }
    out["attention"] = cases

    # plain head
    psd = base["plain_head"]["state_dict"]
    torch.manual_seed(13)
    with quiet:
        pdec = ref_modules.ClassificationDecoder(16, 24, 24, 11)
    pdec.load_state_dict(psd)
    pdec.train()
    for m in pdec.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    lat = torch.randn(3, 16, 24, 24)
    tgt = (torch.rand(3, 11) < 0.3).float()
    with quiet:
        logits = pdec(lat)
    loss = ref_losses.FocalLoss(1.0, 2.0)(logits, tgt)
    loss.backward()
    out["plain"] = {"latent": lat, "targets": tgt, "logits": logits.detach().clone(), "loss": loss.detach().clone(),
                    "grads": {k: digest(p.grad) for k, p in pdec.named_parameters()},
                    "param_order": [k for k, _ in pdec.named_parameters()]}

    # --use_cross_attention head (modules.py:388-395, :450-459), eval mode: the shared parameters are those of
    # head_golden.pt, the cross-attention / query_generator parameters are stored here
    torch.manual_seed(16)
    with quiet:
        xdec = ref_modules.create_attention_decoder(16, 24, 40, 11, attention_config={"use_cross_attention": True})
    xdec.load_state_dict(sd, strict=False)
    xdec.eval()
    lat = torch.randn(3, 16, 24, 40) * 0.5 + 0.1
    with torch.no_grad(), quiet:
        xl = xdec(lat)
    out["cross_attention_head"] = {
        "extra_state_dict": {k: v.clone() for k, v in xdec.state_dict().items()
                             if k.startswith(("cross_attention.", "query_generator."))},
        "latent": lat, "logits": xl,
    }
    # the same head in train mode (dropout off): gradients of the reference's own autograd graph
    xdec.train()
    for m_ in xdec.modules():
        if isinstance(m_, torch.nn.Dropout):
            m_.p = 0.0
    tgt = (torch.rand(3, 11) < 0.3).float()
    with quiet:
        xlogits = xdec(lat)
    xloss = ref_losses.FocalLoss(1.0, 2.0)(xlogits, tgt)
    xloss.backward()
    out["cross_attention_head"]["train"] = {
        "targets": tgt, "logits": xlogits.detach().clone(), "loss": xloss.detach().clone(),
        "grads": {k: digest(p.grad) for k, p in xdec.named_parameters()},
        "param_order": [k for k, _ in xdec.named_parameters()],
    }

    # ClassBalancedLoss (improved_losses.py:58-72) value and gradient
    torch.manual_seed(15)
    x = (torch.randn(6, 11) * 2).requires_grad_(True)
    y = (torch.rand(6, 11) < 0.3).float()
    spc = np.array([120.0, 3.0, 45.0, 1.0, 800.0, 17.0, 5.0, 260.0, 9.0, 33.0, 2.0])
    cb = ref_losses.ClassBalancedLoss()(x, y, spc)
    cb.backward()
    out["class_balanced"] = {"logits": x.detach().clone(), "targets": y, "samples_per_class": spc,
                             "loss": cb.detach().clone(), "grad": x.grad.clone()}

    # plain head without adaptive pooling (modules.py:316-317): a 16x4x4 latent flattens to the same 256 inputs, so
    # the parameters are those of head_golden.pt's plain head
    torch.manual_seed(17)
    with quiet:
        fdec = ref_modules.ClassificationDecoder(16, 4, 4, 11, use_adaptive_pooling=False)
    fdec.load_state_dict(psd)
    lat = torch.randn(5, 16, 4, 4)
    tgt = (torch.rand(5, 11) < 0.3).float()
    fdec.eval()
    with torch.no_grad(), quiet:
        flogits_eval = fdec(lat)
    fdec.train()
    for m_ in fdec.modules():
        if isinstance(m_, torch.nn.Dropout):
            m_.p = 0.0
    with quiet:
        flogits = fdec(lat)
    floss = ref_losses.FocalLoss(1.0, 2.0)(flogits, tgt)
    floss.backward()
    out["plain_flat"] = {"latent": lat, "targets": tgt, "logits_eval": flogits_eval, "logits": flogits.detach().clone(),
                         "loss": floss.detach().clone(),
                         "grads": {k: digest(p.grad) for k, p in fdec.named_parameters()}}

    # clip_grad_norm_ + AdamW on a flat tensor, three steps
    torch.manual_seed(14)
    p = torch.nn.Parameter(torch.randn(1000))
    p0 = p.detach().clone()
    opt = torch.optim.AdamW([p], lr=1e-3, weight_decay=1e-2)
    gs, ps_, norms = [], [], []
    for step in range(3):
        g = torch.randn(1000) * (0.01 if step == 1 else 0.1)   # step 1 stays below the clip threshold
        p.grad = g.clone()
        norms.append(torch.nn.utils.clip_grad_norm_([p], 1.0).clone())
        opt.step()
        gs.append(g)
        ps_.append(p.detach().clone())
    out["adamw"] = {"p0": p0, "grads": gs, "params": ps_, "norms": norms, "lr": 1e-3, "wd": 1e-2, "max_norm": 1.0}

    torch.save(out, os.path.join(HERE, "head_train_golden.pt"))
    print("wrote head_train_golden.pt", os.path.getsize(os.path.join(HERE, "head_train_golden.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
