import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from vae_tagger_b200 import _native as N
ctx = N.get_context(0)
def rel(a, b): return ((a.cpu() - b).norm() / b.norm().clamp_min(1e-30)).item()
g = torch.Generator().manual_seed(0)
n, cin, cout, h, w, k = [int(v) for v in sys.argv[1:7]]
x = torch.randn(n, cin, h, w, generator=g, requires_grad=True)
wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).requires_grad_()
b = torch.randn(cout, generator=g, requires_grad=True)
go = torch.randn(n, cout, h, w, generator=g)
F.conv2d(x, wt, b, padding=k // 2).backward(go)
for name, want in (("wgrad", (False, True, False)),):
    try:
        r = ctx.op_conv2d_backward(x, wt, go, precision=N.PREC_F16, want=want)
        torch.cuda.synchronize()
        ref = {"bias": b.grad, "dgrad": x.grad, "wgrad": wt.grad}[name]
        got = [t for t in r if t is not None][0]
        print(name, "rel", rel(got, ref), flush=True)
        gc = got.cpu()
        for t in range(k * k):
            print(" tap", t, rel(gc[:, :, t // k, t % k], ref[:, :, t // k, t % k]), flush=True)
    except Exception as e:
        print(name, "FAILED", str(e)[:300], flush=True)
        break
