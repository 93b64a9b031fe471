import sys, torch
sys.path.insert(0, "/root/repo")
from oracle.encoder import make_oracle_vae, oracle_wrapper_encode, synthetic_images
from vae_tagger_b200 import diffusers_vae_loader as L
oracle = make_oracle_vae(0)
vae = L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())
vae.load_state_dict(oracle.state_dict(), strict=False)
wrap = L.DiffusersVAEWrapper(vae).cuda().eval()
def rel(a, b): return ((a - b).norm() / b.norm()).item()
for (B, H, W) in [(2, 72, 88), (1, 520, 776), (1, 8, 8), (1, 16, 24), (3, 40, 8), (1, 264, 1000)]:
    x = synthetic_images(B, H, W)
    with torch.no_grad():
        ref = oracle_wrapper_encode(oracle, x)
    out = []
    for prec in ("fp32", "bf16"):
        wrap.vae.precision = prec
        try:
            got = wrap.encode(x.cuda()).cpu()
            out.append(f"{prec} {rel(got, ref):.2e}")
        except Exception as e:
            out.append(f"{prec} ERR {str(e)[:120]}")
    print((B, H, W), ref.shape, out, flush=True)
