"""Golden values + autograd gradients of the reference's OWN VAE fine-tuning losses (improved_losses.py:6-37, 74-125,
127-339), run in the build container only:

    python tests/golden/make_loss_golden.py      ->  tests/golden/loss_golden.pt

Cases: ImprovedTripletLoss / ContrastiveLoss (cosine and euclidean, with and without label weights, rows on both
sides of the hinge and of the 0.3 Jaccard threshold), AdaptiveLossWeights, SimplifiedCombinedLoss (triplet and
contrastive variants) and CombinedLoss (fixed and adaptive weights) -- inputs, loss values and every input gradient.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_stub_diffusers  # noqa: E402


class Posterior:
    """mean / logvar holder with diffusers' kl() (what CombinedLoss calls)."""

    def __init__(self, mean, logvar):
        self.mean, self.logvar = mean, logvar

    def kl(self):
        return 0.5 * torch.sum(self.mean ** 2 + self.logvar.exp() - 1.0 - self.logvar, dim=[1, 2, 3])


def main():
    install_stub_diffusers()
    sys.path.insert(0, REF)
    import improved_losses as R  # noqa: E402

    g = torch.Generator().manual_seed(11)
    out = {"triplet": [], "contrastive": [], "adaptive": [], "simplified": [], "combined": []}

    def rnd(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    def labels(B, T, p=0.3):
        return (torch.rand(B, T, generator=g) < p).float()

    # ---- triplet
    for sim in ("cosine", "euclidean"):
        for B, D, T, margin, with_labels in ((6, 1024, 11, 1.0, True), (5, 4100, 37, 0.2, False), (3, 16 * 16 * 16, 5, 0.5, True)):
            a = rnd(B, D).requires_grad_()
            p = (a.detach() + rnd(B, D, scale=0.7)).requires_grad_()
            n = rnd(B, D).requires_grad_()
            with torch.no_grad():   # some rows far inside the margin (inactive hinge)
                n[0] = -a[0] * (1.0 if sim == "cosine" else 3.0)
            la, lp = (labels(B, T), labels(B, T)) if with_labels else (None, None)
            if with_labels:
                la[1] = 0   # an anchor without labels: weight 1 + 0.5 * 0 / 1e-8
            loss = R.ImprovedTripletLoss(margin=margin, similarity_type=sim)(a, p, n, la, lp)
            loss.backward()
            out["triplet"].append(dict(sim=sim, margin=margin, a=a.detach(), p=p.detach(), n=n.detach(), la=la, lp=lp,
                                       loss=loss.detach(), ga=a.grad, gp=p.grad, gn=n.grad))
    # ---- contrastive
    for sim in ("cosine", "euclidean"):
        for B, D, T, margin in ((8, 1024, 11, 1.0), (4, 4100, 37, 2.5)):
            e1 = rnd(B, D, scale=0.05 if sim == "euclidean" else 1.0).requires_grad_()
            e2 = (e1.detach() + rnd(B, D, scale=0.02 if sim == "euclidean" else 0.7)).requires_grad_()
            l1 = labels(B, T, 0.5)
            l2 = l1.clone()
            l2[B // 2:] = labels(B - B // 2, T, 0.5)   # first half: identical label sets (similar), rest: random
            loss = R.ContrastiveLoss(margin=margin, similarity_type=sim)(e1, e2, l1, l2)
            loss.backward()
            out["contrastive"].append(dict(sim=sim, margin=margin, a=e1.detach(), p=e2.detach(), la=l1, lp=l2,
                                           loss=loss.detach(), ga=e1.grad, gp=e2.grad))
    # ---- adaptive weights
    for temp in (1.0, 0.5):
        m = R.AdaptiveLossWeights(num_losses=4, temperature=temp)
        with torch.no_grad():
            m.log_weights.copy_(rnd(4))
        ls = [rnd(1).abs().squeeze().requires_grad_() for _ in range(4)]
        total, w = m(ls)
        total.backward()
        out["adaptive"].append(dict(temp=temp, log_w=m.log_weights.detach().clone(), losses=torch.stack([l.detach() for l in ls]),
                                    total=total.detach(), weights=w.detach(), g_log_w=m.log_weights.grad.clone(),
                                    g_losses=torch.stack([l.grad for l in ls])))
    # ---- SimplifiedCombinedLoss
    for use_contrastive in (False, True):
        B, T = 4, 11
        z = [rnd(B, 16, 8, 8).requires_grad_() for _ in range(3)]
        logits = rnd(B, T).requires_grad_()
        y, lp = labels(B, T), labels(B, T)
        fn = R.SimplifiedCombinedLoss(use_contrastive=use_contrastive, contrastive_weight=0.7 if use_contrastive else 0.0)
        d = fn(z[0], z[1], z[2], logits, y, y, lp)
        d["total_loss"].backward()
        out["simplified"].append(dict(use_contrastive=use_contrastive, z=[t.detach() for t in z], logits=logits.detach(), y=y, lp=lp,
                                      result={k: v.detach() for k, v in d.items()},
                                      gz=[t.grad for t in z], glogits=logits.grad))
    # ---- CombinedLoss
    for adaptive in (False, True):
        B, T = 3, 11
        recon, target = rnd(B, 3, 32, 32).requires_grad_(), rnd(B, 3, 32, 32)
        means = [rnd(B, 16, 4, 4).requires_grad_() for _ in range(3)]
        logvars = [rnd(B, 16, 4, 4, scale=0.5).requires_grad_() for _ in range(3)]
        z = [rnd(B, 16, 4, 4).requires_grad_() for _ in range(3)]
        logits = rnd(B, T).requires_grad_()
        y, lp = labels(B, T), labels(B, T)
        fn = R.CombinedLoss(use_adaptive_weights=adaptive)
        if adaptive:
            with torch.no_grad():
                fn.adaptive_weights.log_weights.copy_(rnd(4) * 0.5)
        d = fn(recon, target, *[Posterior(m, lv) for m, lv in zip(means, logvars)], z[0], z[1], z[2], logits, y, y, lp)
        d["total_loss"].backward()
        out["combined"].append(dict(
            adaptive=adaptive, recon=recon.detach(), target=target, means=[t.detach() for t in means],
            logvars=[t.detach() for t in logvars], z=[t.detach() for t in z], logits=logits.detach(), y=y, lp=lp,
            log_w=fn.adaptive_weights.log_weights.detach().clone() if adaptive else None,
            result={k: v.detach() for k, v in d.items()}, g_recon=recon.grad, g_means=[t.grad for t in means],
            g_logvars=[t.grad for t in logvars], gz=[t.grad for t in z], glogits=logits.grad,
            g_log_w=fn.adaptive_weights.log_weights.grad.clone() if adaptive else None))
    path = os.path.join(HERE, "loss_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
