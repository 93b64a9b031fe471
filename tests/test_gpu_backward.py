"""Encoder backward building blocks (the reference fine-tunes the VAE through autograd in
train_vae.py:124-186 / train_full.py:201-256): data / weight / bias gradients of the convs (tcgen05 implicit GEMMs:
the data gradient is the forward kernel on flipped weights, the weight gradient a split-K GEMM over the pixels),
GroupNorm+SiLU backward, and a whole ResnetBlock2D, against ``torch.autograd`` on the oracle's own modules (CPU fp32).

Tolerances (relative L2 of each gradient tensor): fp32 verification mode <= 1e-4; 16-bit mode <= 2e-2 -- gradients
and the re-laid-out operands are bf16 (8-bit mantissa: the range of a gradient is unbounded and the reference has no
loss scaling), accumulation fp32."""
import pytest
import torch
import torch.nn.functional as F

from oracle.encoder import OracleResnetBlock2D
from vae_tagger_b200 import _native as N

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
B16_TOL = 2e-2


def rel(a, b):
    return ((a.cpu() - b).norm() / b.norm().clamp_min(1e-30)).item()


def tol(prec):
    return FP32_TOL if prec == N.PREC_FP32 else B16_TOL


@pytest.mark.parametrize("prec", [N.PREC_FP32, N.PREC_F16])
@pytest.mark.parametrize("n,cin,cout,h,w,k", [
    (1, 128, 128, 16, 16, 3), (2, 128, 256, 24, 40, 3), (1, 256, 256, 33, 17, 3), (2, 512, 512, 8, 8, 3),
    (1, 256, 128, 20, 12, 3), (2, 128, 256, 16, 24, 1), (1, 512, 512, 64, 64, 3),
])
def test_conv2d_backward(ctx, prec, n, cin, cout, h, w, k):
    g = torch.Generator().manual_seed(n + cin + cout + h + w + k)
    x = torch.randn(n, cin, h, w, generator=g, requires_grad=True)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).requires_grad_()
    b = torch.randn(cout, generator=g, requires_grad=True)
    go = torch.randn(n, cout, h, w, generator=g)
    F.conv2d(x, wt, b, padding=k // 2).backward(go)
    gx, gw, gb = ctx.op_conv2d_backward(x, wt, go, precision=prec)
    assert rel(gx, x.grad) < tol(prec), ("grad_x", rel(gx, x.grad))
    assert rel(gw, wt.grad) < tol(prec), ("grad_w", rel(gw, wt.grad))
    assert rel(gb, b.grad) < tol(prec), ("grad_b", rel(gb, b.grad))


def test_conv2d_backward_is_bit_reproducible(ctx):
    """Split-K partial tiles are added in index order (no atomics): two runs give identical bits."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 128, 48, 40, generator=g)
    wt = torch.randn(256, 128, 3, 3, generator=g) / 34.0
    go = torch.randn(2, 256, 48, 40, generator=g)
    a = ctx.op_conv2d_backward(x, wt, go)
    b = ctx.op_conv2d_backward(x, wt, go)
    assert all(torch.equal(p, q) for p, q in zip(a, b))


@pytest.mark.parametrize("prec", [N.PREC_FP32, N.PREC_F16])
@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("shape", [(2, 128, 16, 24), (1, 256, 9, 7), (3, 512, 12, 4), (1, 128, 96, 96)])
def test_group_norm_backward(ctx, prec, silu, shape):
    g = torch.Generator().manual_seed(shape[1] + shape[2])
    x = (torch.randn(*shape, generator=g) * 1.5 + 0.3).requires_grad_()
    gamma = (torch.randn(shape[1], generator=g) * 0.5 + 1.0).requires_grad_()
    beta = (torch.randn(shape[1], generator=g) * 0.3).requires_grad_()
    gy = torch.randn(*shape, generator=g)
    y = F.group_norm(x, 32, gamma, beta, eps=1e-6)
    (F.silu(y) if silu else y).backward(gy)
    gx, gg, gb = ctx.op_group_norm_backward(x, gamma, beta, gy, silu=silu, precision=prec)
    assert rel(gx, x.grad) < tol(prec), ("grad_x", rel(gx, x.grad))
    assert rel(gg, gamma.grad) < tol(prec), ("grad_gamma", rel(gg, gamma.grad))
    assert rel(gb, beta.grad) < tol(prec), ("grad_beta", rel(gb, beta.grad))


@pytest.mark.parametrize("prec", [N.PREC_FP32, N.PREC_F16])
@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 128, 128, 32, 32), (1, 128, 256, 24, 40), (1, 256, 512, 16, 16),
                                            (2, 512, 512, 16, 8), (1, 128, 128, 64, 64)])
def test_resnet_block_backward(ctx, prec, n, cin, cout, h, w):
    """One ResnetBlock2D of the encoder: d out / d x and every parameter gradient vs autograd on the oracle block."""
    torch.manual_seed(cin + cout + h)
    blk = OracleResnetBlock2D(cin, cout, 32)
    with torch.no_grad():   # non-trivial affine parameters
        for m in (blk.norm1, blk.norm2):
            m.weight.add_(torch.randn_like(m.weight) * 0.3)
            m.bias.add_(torch.randn_like(m.bias) * 0.3)
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(n, cin, h, w, generator=g) * 1.2 + 0.2).requires_grad_()
    go = torch.randn(n, cout, h, w, generator=g)
    blk(x).backward(go)
    gx, grads = ctx.op_resnet_block_backward(x, blk.state_dict(), go, precision=prec)
    errs = {"x": rel(gx, x.grad)}
    for k, p in blk.named_parameters():
        errs[k] = rel(grads[k], p.grad)
    assert set(grads) == {k for k, _ in blk.named_parameters()}
    assert max(errs.values()) < tol(prec), errs


# ---------------------------------------------------------------------------------------------------------
# The whole encoder: training forward + backward (vt_encoder_train_forward / vt_encoder_backward) through the
# reference-facing module (AutoencoderKL.encode in train() mode) against autograd on the oracle encoder.
def _encoder_pair(seed=0):
    from oracle.encoder import make_oracle_vae
    from vae_tagger_b200 import diffusers_vae_loader as L
    oracle = make_oracle_vae(seed=seed)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    missing, unexpected = vae.load_state_dict(oracle.state_dict(), strict=False)
    assert not missing and not unexpected
    for m in (oracle, vae):
        for q in m.parameters():
            q.requires_grad_(True)
    return oracle, vae.cuda()


@pytest.mark.parametrize("prec,n,h,w", [("fp32", 2, 64, 64), ("bf16", 2, 64, 64), ("bf16", 1, 96, 160), ("fp32", 1, 32, 48),
                                        ("bf16", 2, 256, 256)])   # 256^2: many pixel patches per CTA, several K splits
def test_encoder_backward_matches_autograd(prec, n, h, w):
    from oracle.encoder import structured_images, synthetic_images
    oracle, vae = _encoder_pair()
    x = torch.cat([synthetic_images(1, h, w), structured_images(n - 1, h, w)]) if n > 1 else structured_images(1, h, w)
    g = torch.Generator().manual_seed(3)
    gm = torch.randn(n, 16, h // 8, w // 8, generator=g)
    gl = torch.randn(n, 16, h // 8, w // 8, generator=g) * 0.3
    # oracle: autograd through the restated encoder
    oracle.train()
    for p in oracle.parameters():
        p.grad = None
    dist = oracle.encode(x).latent_dist
    ((dist.mean * gm).sum() + (dist.logvar * gl).sum()).backward()
    ref = {k: p.grad for k, p in oracle.encoder.named_parameters()}
    # native
    vae.train()
    vae.precision = prec
    post = vae.encode(x.cuda()).latent_dist
    assert post.mean.requires_grad and post.logvar.requires_grad
    assert rel(post.mean.detach(), dist.mean.detach()) < (1e-4 if prec == "fp32" else 1e-2)
    ((post.mean * gm.cuda()).sum() + (post.logvar * gl.cuda()).sum()).backward()
    got = {k: p.grad for k, p in vae.encoder.named_parameters()}
    assert set(got) == set(ref) and len(ref) == 106
    bar = FP32_TOL if prec == "fp32" else 3e-2
    # d loss / d to_k.bias is analytically ZERO (q . b_k shifts every score of a query by the same amount and softmax
    # ignores it): both sides hold round-off only -- compare against the scale of the sibling gradient instead
    kb = "mid_block.attentions.0.to_k.bias"
    scale = ref["mid_block.attentions.0.to_q.bias"].norm().item()
    assert ref[kb].norm().item() < 1e-4 * scale and got[kb].cpu().norm().item() < (1e-4 if prec == "fp32" else 2e-2) * scale
    errs = {k: rel(got[k], ref[k]) for k in ref if k != kb}
    worst = max(errs, key=errs.get)
    print(f"encoder backward {prec} {n}x{h}x{w}: worst {worst} {errs[worst]:.3e}, "
          f"median {sorted(errs.values())[len(errs) // 2]:.3e}", file=__import__("sys").stderr)
    assert errs[worst] < bar, {k: v for k, v in errs.items() if v >= bar}


def test_three_forwards_then_one_backward():
    """train_full.py:210-212 runs the VAE on anchor, positive and negative before ONE backward: every forward keeps its
    own tape, gradients add up like autograd's, and an eval-mode / no-grad encode leaves no graph."""
    from oracle.encoder import structured_images
    oracle, vae = _encoder_pair()
    vae.precision = "fp32"
    xs = [structured_images(1, 32, 32, seed=s) for s in (1, 2, 3)]
    oracle.train()
    for p in oracle.parameters():
        p.grad = None
    sum(oracle.encode(x).latent_dist.mean.square().sum() for x in xs).backward()
    vae.train()
    sum(vae.encode(x.cuda()).latent_dist.mean.square().sum() for x in xs).backward()
    for k, p in oracle.encoder.named_parameters():
        if k == "mid_block.attentions.0.to_k.bias":   # analytically zero (see above)
            continue
        assert rel(dict(vae.encoder.named_parameters())[k].grad, p.grad) < FP32_TOL, k
    vae.eval()
    assert not vae.encode(xs[0].cuda()).latent_dist.mean.requires_grad
    vae.train()
    with torch.no_grad():
        assert not vae.encode(xs[0].cuda()).latent_dist.mean.requires_grad


# ---------------------------------------------------------------------------------------------------------
# The decoder half: training forward + backward (vt_decoder_train_forward / vt_decoder_backward) behind
# AutoencoderKL.decode in train() mode, against autograd on the oracle decoder -- every parameter gradient and
# d loss / d latent (the path the reconstruction MSE takes back into the encoder, train_vae.py:124-186).
@pytest.mark.parametrize("prec,n,lh,lw,scale_shift", [("fp32", 2, 8, 8, False), ("bf16", 2, 8, 8, True), ("bf16", 1, 12, 20, False),
                                                      ("fp32", 1, 4, 6, True), ("bf16", 1, 32, 32, True)])
def test_decoder_backward_matches_autograd(prec, n, lh, lw, scale_shift):
    from oracle.decoder import make_oracle_decoder, oracle_wrapper_decode
    from vae_tagger_b200 import diffusers_vae_loader as L
    dec = make_oracle_decoder(seed=0)
    for q in dec.parameters():
        q.requires_grad_(True)
    vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
    vae.enable_decoder()
    missing, unexpected = vae.load_state_dict({"decoder." + k: v for k, v in dec.state_dict().items()}, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing)
    vae = vae.cuda().train()
    for q in vae.decoder.parameters():
        q.requires_grad_(True)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(n, 16, lh, lw, generator=g)
    gi = torch.randn(n, 3, 8 * lh, 8 * lw, generator=g)
    zr = z.clone().requires_grad_()
    img_ref = oracle_wrapper_decode(dec, zr) if scale_shift else dec(zr)
    (img_ref * gi).sum().backward()
    ref = {k: p.grad for k, p in dec.named_parameters()}
    vae.precision = prec
    zc = z.cuda().requires_grad_()
    img = vae.decode(zc, apply_scale_shift=scale_shift).sample
    assert img.requires_grad
    assert rel(img.detach(), img_ref.detach()) < (1e-4 if prec == "fp32" else 2e-2)
    (img * gi.cuda()).sum().backward()
    got = {k: p.grad for k, p in vae.decoder.named_parameters()}
    assert set(got) == set(ref) and len(ref) == 138
    bar = FP32_TOL if prec == "fp32" else 3e-2
    kb = "mid_block.attentions.0.to_k.bias"      # analytically zero gradient (see the encoder test)
    scale = ref["mid_block.attentions.0.to_q.bias"].norm().item()
    assert got[kb].cpu().norm().item() < (1e-4 if prec == "fp32" else 2e-2) * scale
    errs = {k: rel(got[k], ref[k]) for k in ref if k != kb}
    errs["latent"] = rel(zc.grad, zr.grad)
    worst = max(errs, key=errs.get)
    print(f"decoder backward {prec} {n}x{lh}x{lw}: worst {worst} {errs[worst]:.3e}, latent {errs['latent']:.3e}, "
          f"median {sorted(errs.values())[len(errs) // 2]:.3e}", file=__import__("sys").stderr)
    assert errs[worst] < bar, {k: v for k, v in errs.items() if v >= bar}


def test_forwards_without_backward_do_not_leak_tape_slots():
    """A train()-mode forward whose graph is dropped (a loss that is only logged) gives its tape slot back: more such
    forwards than the context has slots must keep working, and a later forward + backward is still correct."""
    from oracle.encoder import structured_images
    from vae_tagger_b200 import _native as NV
    oracle, vae = _encoder_pair()
    vae.precision = "fp32"
    vae.train()
    x = structured_images(1, 32, 32, seed=9).cuda()
    for _ in range(NV.MAX_TAPES + 3):
        post = vae.encode(x).latent_dist
        assert post.mean.requires_grad
        del post
    for p in oracle.parameters():
        p.grad = None
    oracle.train()
    oracle.encode(x.cpu()).latent_dist.mean.square().sum().backward()
    vae.encode(x).latent_dist.mean.square().sum().backward()
    k = "down_blocks.0.resnets.0.conv1.weight"
    assert rel(dict(vae.encoder.named_parameters())[k].grad, dict(oracle.encoder.named_parameters())[k].grad) < FP32_TOL


def test_weight_gradient_box_form_agrees():
    """VT_B200_NO_WGRAD_HALO=1 sends the 3x3 stride-1 weight gradients through the shifted-box kernel (the one the
    stride-2 and 1x1 convs always use) instead of the halo-tile kernel; both must meet the same bars.  The switch is
    read once per process, so the conv / ResnetBlock cases run again in a fresh interpreter."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e["VT_B200_NO_WGRAD_HALO"] = "1"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_backward.py"), "-q", "-m", "gpu", "-x",
                          "-k", "test_conv2d_backward or test_resnet_block_backward"], env=e, cwd=root, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
