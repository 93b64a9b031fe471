"""VAE fine-tuning losses against values AND autograd gradients of the reference's own
``improved_losses.py`` (``tests/golden/loss_golden.pt``, made by ``make_loss_golden.py`` from /root/reference):
ImprovedTripletLoss (:74-109), ContrastiveLoss (:6-37), AdaptiveLossWeights (:111-125), SimplifiedCombinedLoss
(:127-232), CombinedLoss (:234-339).  fp32 arithmetic: values to 1e-5 relative, gradients to 1e-4 relative L2."""
import os

import pytest
import torch

from vae_tagger_b200 import improved_losses as L

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "loss_golden.pt"), map_location="cpu", weights_only=False)


def close(got, want, tol=1e-5):
    got, want = got.detach().cpu().double(), want.double()
    return (got - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())


def grad_ok(got, want, tol=1e-4):
    got, want = got.detach().cpu().double(), want.double()
    if want.norm().item() == 0.0:
        return got.abs().max().item() <= 1e-7
    return ((got - want).norm() / want.norm()).item() <= tol


def cu(t):
    return None if t is None else t.cuda()


def test_triplet_loss(gold):
    assert len(gold["triplet"]) == 6
    for c in gold["triplet"]:
        a, p, n = (c[k].cuda().requires_grad_() for k in ("a", "p", "n"))
        loss = L.ImprovedTripletLoss(margin=c["margin"], similarity_type=c["sim"])(a, p, n, cu(c["la"]), cu(c["lp"]))
        loss.backward()
        assert close(loss, c["loss"]), (c["sim"], loss.item(), c["loss"].item())
        assert grad_ok(a.grad, c["ga"]) and grad_ok(p.grad, c["gp"]) and grad_ok(n.grad, c["gn"]), c["sim"]


def test_contrastive_loss(gold):
    assert len(gold["contrastive"]) == 4
    for c in gold["contrastive"]:
        a, p = (c[k].cuda().requires_grad_() for k in ("a", "p"))
        loss = L.ContrastiveLoss(margin=c["margin"], similarity_type=c["sim"])(a, p, cu(c["la"]), cu(c["lp"]))
        loss.backward()
        assert close(loss, c["loss"]), (c["sim"], loss.item(), c["loss"].item())
        assert grad_ok(a.grad, c["ga"]) and grad_ok(p.grad, c["gp"]), c["sim"]


def test_adaptive_loss_weights(gold):
    for c in gold["adaptive"]:
        m = L.AdaptiveLossWeights(num_losses=4, temperature=c["temp"]).cuda()
        with torch.no_grad():
            m.log_weights.copy_(c["log_w"])
        ls = [v.clone().cuda().requires_grad_() for v in c["losses"]]
        total, w = m(ls)
        total.backward()
        assert close(total, c["total"]) and close(w, c["weights"])
        assert grad_ok(m.log_weights.grad, c["g_log_w"])
        assert grad_ok(torch.stack([v.grad for v in ls]), c["g_losses"])


def test_simplified_combined_loss(gold):
    for c in gold["simplified"]:
        z = [t.cuda().requires_grad_() for t in c["z"]]
        logits = c["logits"].cuda().requires_grad_()
        fn = L.SimplifiedCombinedLoss(use_contrastive=c["use_contrastive"],
                                      contrastive_weight=0.7 if c["use_contrastive"] else 0.0)
        d = fn(z[0], z[1], z[2], logits, c["y"].cuda(), c["y"].cuda(), c["lp"].cuda())
        d["total_loss"].backward()
        assert set(d) == set(c["result"])
        for k, v in c["result"].items():
            assert close(d[k], v), k
        for got, want in zip(z, c["gz"]):
            if want is None:
                assert got.grad is None
            else:
                assert grad_ok(got.grad, want)
        assert grad_ok(logits.grad, c["glogits"])


def test_combined_loss(gold):
    class Posterior:
        def __init__(self, mean, logvar):
            self.mean, self.logvar = mean, logvar

        def kl(self):
            return 0.5 * torch.sum(self.mean ** 2 + self.logvar.exp() - 1.0 - self.logvar, dim=[1, 2, 3])

    for c in gold["combined"]:
        recon = c["recon"].cuda().requires_grad_()
        means = [t.cuda().requires_grad_() for t in c["means"]]
        logvars = [t.cuda().requires_grad_() for t in c["logvars"]]
        z = [t.cuda().requires_grad_() for t in c["z"]]
        logits = c["logits"].cuda().requires_grad_()
        fn = L.CombinedLoss(use_adaptive_weights=c["adaptive"]).cuda()
        if c["adaptive"]:
            with torch.no_grad():
                fn.adaptive_weights.log_weights.copy_(c["log_w"])
        d = fn(recon, c["target"].cuda(), *[Posterior(m, lv) for m, lv in zip(means, logvars)], z[0], z[1], z[2], logits,
               c["y"].cuda(), c["y"].cuda(), c["lp"].cuda())
        d["total_loss"].backward()
        assert set(d) == set(c["result"])
        for k, v in c["result"].items():
            assert close(d[k], v), k
        assert grad_ok(recon.grad, c["g_recon"]) and grad_ok(logits.grad, c["glogits"])
        for got, want in zip(means + logvars + z, c["g_means"] + c["g_logvars"] + c["gz"]):
            assert grad_ok(got.grad, want)
        if c["adaptive"]:
            assert grad_ok(fn.adaptive_weights.log_weights.grad, c["g_log_w"])


def test_losses_refuse_cpu_tensors():
    from vae_tagger_b200._native import NativeError
    with pytest.raises(NativeError):
        L.ImprovedTripletLoss()(torch.zeros(2, 8), torch.zeros(2, 8), torch.zeros(2, 8))
    with pytest.raises(NativeError):
        L.ContrastiveLoss()(torch.zeros(2, 8), torch.zeros(2, 8), torch.zeros(2, 3), torch.zeros(2, 3))
