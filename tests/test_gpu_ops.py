"""Single-kernel parity: each CUDA kernel family through the C-ABI op entry points against a
plain PyTorch fp32 reference of the same op (floating-point kernels; tolerances stated per test).

bf16 mode: operands are rounded to bf16, accumulation is fp32 in TMEM -> the reference is computed
from bf16-rounded operands in fp32, leaving only accumulation-order error (<= 2e-3 relative to the
output scale is a loose bound; observed ~1e-6).
"""
import os

import pytest
import torch
import torch.nn.functional as F

from vae_tagger_b200 import _native as N

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def r16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def rh(t):
    return t.to(torch.float16).to(torch.float32)


# storage format of RAW activations (residual stream, shortcut operand, GroupNorm input) in the 16-bit modes:
# fp16 by default, bf16 with VT_B200_RAW_BF16=1 (vt_ctx::raw_f16)
RAW_F16 = os.environ.get("VT_B200_RAW_BF16", "0")[:1] != "1"


def raw(t):
    return rh(t) if RAW_F16 else r16(t)


def opd_round(prec):
    """rounding applied to the main operands by each precision mode of the single-op entry points"""
    return {N.PREC_BF16: r16, N.PREC_F16: rh, N.PREC_FP32: (lambda t: t)}[prec]


@pytest.mark.parametrize("batch,M,Nn,K,b_batched", [
    (1, 128, 128, 64, 0), (1, 256, 256, 512, 0), (2, 192, 1536, 512, 0), (3, 64, 32, 128, 1),
    (2, 320, 448, 1024, 1), (1, 1024, 4096, 512, 1),
])
@pytest.mark.parametrize("prec", [N.PREC_BF16, N.PREC_F16])
def test_gemm_16bit(ctx, batch, M, Nn, K, b_batched, prec):
    g = torch.Generator().manual_seed(M + Nn + K)
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch if b_batched else 1, Nn, K, generator=g)
    bias = torch.randn(Nn, generator=g)
    out = ctx.op_gemm_nt(A, B if b_batched else B[0], bias, alpha=0.5, precision=prec).cpu()
    rr = opd_round(prec)
    ref = 0.5 * torch.matmul(rr(A), rr(B).transpose(1, 2)) + bias
    assert rel(out, ref) < 2e-5, (rel(out, ref), (out - ref).abs().max().item())


@pytest.mark.parametrize("batch,M,Nn,K", [(1, 70, 50, 36), (2, 128, 64, 512)])
def test_gemm_fp32(ctx, batch, M, Nn, K):
    g = torch.Generator().manual_seed(7)
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch, Nn, K, generator=g)
    out = ctx.op_gemm_nt(A, B, None, alpha=1.0, precision=N.PREC_FP32).cpu()
    ref = torch.matmul(A, B.transpose(1, 2))
    assert rel(out, ref) < 2e-6


CONV_CASES = [
    # N, Cin, H, W, Cout, k, stride, residual, shortcut Cs
    (1, 128, 16, 16, 128, 3, 1, False, 0),
    (2, 128, 24, 40, 128, 3, 1, True, 0),
    (1, 128, 32, 32, 256, 3, 1, False, 0),
    (1, 256, 16, 16, 256, 3, 1, False, 128),
    (2, 512, 8, 8, 512, 3, 1, True, 0),
    (1, 256, 16, 24, 512, 3, 1, False, 256),
    (1, 128, 32, 32, 128, 3, 2, False, 0),
    (2, 256, 16, 48, 256, 3, 2, False, 0),
    (1, 512, 16, 16, 32, 3, 1, False, 0),
    (1, 64, 16, 16, 128, 1, 1, False, 0),
    (1, 512, 72, 8, 512, 3, 1, True, 0),
]


def conv_ref(x, w, b, res, scx, scw, stride):
    if stride == 2:
        y = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    else:
        y = F.conv2d(x, w, b, padding=w.shape[-1] // 2)
    if scx is not None:
        y = y + F.conv2d(scx, scw)
    if res is not None:
        y = y + res
    return y


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("prec", [N.PREC_BF16, N.PREC_F16, N.PREC_FP32])
def test_conv(ctx, case, prec):
    n, cin, h, w_, cout, k, stride, use_res, cs = case
    g = torch.Generator().manual_seed(sum(case[:7]))
    x = torch.randn(n, cin, h, w_, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ho, wo = (h, w_) if stride == 1 else (h // 2, w_ // 2)
    res = torch.randn(n, cout, ho, wo, generator=g) if use_res else None
    scx = torch.randn(n, cs, ho, wo, generator=g) if cs else None
    scw = torch.randn(cout, cs, 1, 1, generator=g) / cs ** 0.5 if cs else None
    want_stats = cout in (128, 256, 512)
    r = ctx.op_conv2d(x, w, b, res, scx, scw, stride=stride, precision=prec, want_stats=want_stats)
    out, stats = (r if want_stats else (r, None))
    out = out.cpu()
    if prec != N.PREC_FP32:
        # main operand + weights in the mode's format; residual / shortcut operands are raw 16-bit tensors
        rr = opd_round(prec)
        ref = conv_ref(rr(x), rr(w), b, raw(res) if use_res else None, raw(scx) if cs else None,
                       raw(scw) if cs else None, stride)
        tol = 3e-5
    else:
        ref = conv_ref(x, w, b, res, scx, scw, stride)
        tol = 3e-6
    assert rel(out, ref) < tol, (rel(out, ref), (out - ref).abs().max().item())
    if stats is not None:
        # fused GroupNorm statistics: (sum, sumsq) over (pixels, channels of the group), 32 groups
        grp = ref.double().reshape(n, 32, -1)
        want = torch.stack([grp.sum(-1), (grp * grp).sum(-1)], dim=-1)
        got = stats.cpu()
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-3 * grp.shape[-1] ** 0.5), (got - want).abs().max().item()


@pytest.mark.parametrize("shape", [(2, 128, 16, 24), (1, 256, 8, 8), (3, 512, 12, 4)])
@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("prec", [N.PREC_BF16, N.PREC_F16, N.PREC_FP32])
def test_group_norm(ctx, shape, silu, prec):
    g = torch.Generator().manual_seed(shape[1])
    x = torch.randn(*shape, generator=g) * 2 + 0.7
    gamma = torch.randn(shape[1], generator=g)
    beta = torch.randn(shape[1], generator=g)
    out = ctx.op_group_norm(x, gamma, beta, silu=silu, precision=prec).cpu()
    # raw input storage: the context's raw format in fp16 mode, bf16 in bf16 mode
    xin = x if prec == N.PREC_FP32 else (raw(x) if prec == N.PREC_F16 else r16(x))
    ref = F.group_norm(xin, 32, gamma, beta, eps=1e-6)
    if silu:
        ref = F.silu(ref)
    # the result is rounded to bf16 (half an ulp = 2^-9 relative) or fp16 (2^-12)
    tol = {N.PREC_BF16: 4e-3, N.PREC_F16: 5e-4, N.PREC_FP32: 2e-6}[prec]
    assert rel(out, ref) < tol, rel(out, ref)


@pytest.mark.parametrize("rows,cols", [(5, 64), (3, 1000), (4, 4096), (2, 16384), (2, 20000)])
@pytest.mark.parametrize("prec", [N.PREC_BF16, N.PREC_F16, N.PREC_FP32])
def test_softmax_rows(ctx, rows, cols, prec):
    g = torch.Generator().manual_seed(cols)
    s = torch.randn(rows, cols, generator=g) * 4
    out = ctx.op_softmax_rows(s, precision=prec).cpu()
    ref = torch.softmax(s, dim=-1)
    tol = {N.PREC_BF16: 4e-3, N.PREC_F16: 5e-4, N.PREC_FP32: 2e-6}[prec]
    assert rel(out, ref) < tol, rel(out, ref)


def test_conv3_fused_with_shortcut_slab(ctx):
    """conv2 of a channel-changing ResnetBlock2D: conv3x3(silu(gn(h))) + W_s x, the 1x1 shortcut as extra K slab."""
    for n, cin, cs, h, w_, cout in ((1, 256, 128, 32, 24, 256), (2, 512, 256, 16, 16, 512)):
        g = torch.Generator().manual_seed(cin + cs)
        x = torch.randn(n, cin, h, w_, generator=g) * 1.3 - 0.2
        gamma = torch.randn(cin, generator=g) * 0.5 + 1.0
        beta = torch.randn(cin, generator=g) * 0.3
        w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        b = torch.randn(cout, generator=g)
        scx = torch.randn(n, cs, h, w_, generator=g)
        scw = torch.randn(cout, cs, 1, 1, generator=g) / cs ** 0.5
        out = ctx.op_conv3_fused(x, gamma, beta, w, b, None, sc_x=scx, sc_w=scw).cpu()
        t = F.silu(F.group_norm(raw(x), 32, gamma, beta, eps=1e-6))
        ref = F.conv2d(rh(t), rh(w), b, padding=1) + F.conv2d(raw(scx), raw(scw))
        assert rel(out, ref) < 1e-3, (cin, rel(out, ref))


FUSED_CASES = [
    # N, Cin, H, W, Cout, residual
    (1, 128, 16, 16, 128, False),
    (2, 128, 40, 24, 128, True),
    (1, 128, 64, 64, 128, True),
    (1, 256, 16, 32, 256, False),
    (2, 512, 8, 8, 512, True),
    (1, 128, 32, 16, 256, False),
    (1, 512, 72, 24, 512, True),
]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_conv3_fused_groupnorm_silu(ctx, case):
    """conv3x3(silu(GroupNorm32(x))) with the normalisation fused into the operand path (halo tile, nine
    shifted descriptor views).  Reference: x rounded to the raw storage format, GroupNorm+SiLU in fp32, result
    rounded to fp16 (operand), fp16 weights, fp32 accumulation, residual in the raw storage format."""
    n, cin, h, w_, cout, use_res = case
    g = torch.Generator().manual_seed(sum(case[:5]))
    x = torch.randn(n, cin, h, w_, generator=g) * 1.7 + 0.4
    gamma = torch.randn(cin, generator=g) * 0.5 + 1.0
    beta = torch.randn(cin, generator=g) * 0.3
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(n, cout, h, w_, generator=g) if use_res else None
    out, stats = ctx.op_conv3_fused(x, gamma, beta, w, b, res, want_stats=True)
    out = out.cpu()
    t = F.silu(F.group_norm(raw(x), 32, gamma, beta, eps=1e-6))
    ref = F.conv2d(rh(t), rh(w), b, padding=1)
    if use_res:
        ref = ref + raw(res)
    # the kernel evaluates SiLU as h + h*tanh.approx(h) on half2 (h = t/2 in fp16): operand values differ from
    # the fp32 reference rounded to fp16 by up to ~2 fp16 ulps (2^-10 relative)
    assert rel(out, ref) < 1e-3, (rel(out, ref), (out - ref).abs().max().item())
    # fused statistics = (sum, sumsq) of the kernel's own fp32 output per (image, group)
    grp = out.double().reshape(n, 32, -1)
    want = torch.stack([grp.sum(-1), (grp * grp).sum(-1)], dim=-1)
    got = stats.cpu()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-3 * grp.shape[-1] ** 0.5), (got - want).abs().max().item()


@pytest.mark.parametrize("n,tokens", [(1, 64), (2, 128), (1, 320), (2, 1024), (1, 4160)])
def test_flash_attention_d512(ctx, n, tokens):
    """Fused attention (head_dim 512) against softmax(q k^T / sqrt(512)) v + b_v in fp32 on fp16-rounded
    operands.  Tolerance: P and O are rounded to fp16 inside the kernel (2^-11 relative each)."""
    g = torch.Generator().manual_seed(tokens)
    q = torch.randn(n, tokens, 512, generator=g) * 1.5
    k = torch.randn(n, tokens, 512, generator=g) * 1.5
    v = torch.randn(n, tokens, 512, generator=g)
    bv = torch.randn(512, generator=g)
    # make some rows sharply peaked so the running-maximum rescale path is exercised
    k[:, tokens // 2] = q[:, 0] * 3.0
    out = ctx.op_flash_attention(q, k, v, bv).cpu()
    s_ = torch.matmul(rh(q), rh(k).transpose(1, 2)) / 512 ** 0.5
    ref = torch.matmul(torch.softmax(s_, dim=-1), rh(v)) + bv
    assert rel(out, ref) < 2e-3, (rel(out, ref), (out - ref).abs().max().item())
