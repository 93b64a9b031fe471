"""CPU emulation of the bf16 pipeline's rounding points on top of the fp32 oracle modules
(development aid: predicts the GPU path's error vs the oracle and its sensitivity).

    python tests/emulate_bf16.py [resolution]
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.encoder import make_oracle_vae, oracle_wrapper_encode, synthetic_images  # noqa: E402


def r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def gn_stats(x):  # fp32 pre-rounding tensor -> (mean, rstd) per (n, group)
    n = x.shape[0]
    g = x.double().reshape(n, 32, -1)
    mean = g.mean(-1)
    var = (g * g).mean(-1) - mean * mean
    return mean.float(), (1.0 / torch.sqrt(var.clamp_min(0) + 1e-6)).float()


def gn_apply(xb, stats, norm, silu=True, keep_fp32=False):
    c = xb.shape[1]
    mean, rstd = stats
    cpg = c // 32
    mean = mean.repeat_interleave(cpg, 1)[:, :, None, None]
    rstd = rstd.repeat_interleave(cpg, 1)[:, :, None, None]
    sc = rstd * norm.weight[None, :, None, None]
    sh = norm.bias[None, :, None, None] - mean * sc
    y = xb * sc + sh
    if silu:
        y = y / (1 + torch.exp(-y))
    return y if keep_fp32 else r(y)


def conv(x, m, stride=1):
    w = r(m.weight)
    if stride == 2:
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, m.bias, stride=2)
    return F.conv2d(x, w, m.bias, padding=m.kernel_size[0] // 2)


def resnet(rb, xb, st, res_fp32=False):
    t = gn_apply(xb, st, rb.norm1)
    h32 = conv(t, rb.conv1)
    st_h = gn_stats(h32)
    t2 = gn_apply(r(h32), st_h, rb.norm2)
    o = conv(t2, rb.conv2)
    if rb.conv_shortcut is not None:
        o = o + F.conv2d(r(xb), r(rb.conv_shortcut.weight), rb.conv_shortcut.bias)
    else:
        o = o + xb
    return (o if res_fp32 else r(o)), gn_stats(o)


def emulate(vae, x, res_fp32=False):
    e = vae.encoder
    h32 = conv(r(x), e.conv_in)
    st = gn_stats(h32)
    h = h32 if res_fp32 else r(h32)
    for blk in e.down_blocks:
        for rb in blk.resnets:
            h, st = resnet(rb, h, st, res_fp32)
        if blk.downsamplers is not None:
            o = conv(r(h), blk.downsamplers[0].conv, stride=2)
            st = gn_stats(o)
            h = o if res_fp32 else r(o)
    h, st = resnet(e.mid_block.resnets[0], h, st, res_fp32)
    a = e.mid_block.attentions[0]
    b, c, hh, ww = h.shape
    t = gn_apply(h, st, a.group_norm, silu=False).reshape(b, c, hh * ww).transpose(1, 2)
    q = r(F.linear(t, r(a.to_q.weight), a.to_q.bias))
    k = r(F.linear(t, r(a.to_k.weight), a.to_k.bias))
    v = r(F.linear(t, r(a.to_v.weight)))
    s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(c)
    p = r(torch.softmax(s, -1))
    o = r(torch.matmul(p, v) + a.to_v.bias)
    o = F.linear(o, r(a.to_out[0].weight), a.to_out[0].bias).transpose(1, 2).reshape(b, c, hh, ww) + h
    st = gn_stats(o)
    h = o if res_fp32 else r(o)
    h, st = resnet(e.mid_block.resnets[1], h, st, res_fp32)
    t = gn_apply(h, st, e.conv_norm_out)
    mom = conv(t, e.conv_out)
    return mom[:, :16] * 0.3611 + 0.1159


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    vae = make_oracle_vae(0)
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    x = synthetic_images(2, R, R)
    ref = oracle_wrapper_encode(vae, x)
    a = emulate(vae, x)
    print("emulated bf16 vs fp32 oracle      :", rel(a, ref))
    b = emulate(vae, x * (1 + 1e-7))
    print("emulated bf16, 1e-7 perturbed input:", rel(b, a))
    c = emulate(vae, x, res_fp32=True)
    print("emulated bf16 with fp32 residual stream vs oracle:", rel(c, ref))
    ref2 = oracle_wrapper_encode(vae, x * (1 + 1e-7))
    print("fp32 oracle sensitivity to the 1e-7 perturbation  :", rel(ref2, ref))
