"""GPU preprocessing parity: ``vt_resize_u8`` against Pillow itself and the oracle's
restatement of it -- BIT-EXACT (uint8 image arithmetic) -- and the bucket batcher end to end against the
reference's host transform (SmartResize / Resize -> ToTensor -> Normalize) feeding the same encoder."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import resample as R
from vae_tagger_b200 import _native
from vae_tagger_b200 import modules as M
from vae_tagger_b200.preprocess import BucketBatcher, gpu_smart_resize, gpu_square_resize

pytestmark = pytest.mark.gpu


def rnd_img(rng, w, h):
    # smooth structure + noise so that the negative Lanczos lobes and the clipping are both exercised
    y, x = np.mgrid[0:h, 0:w]
    base = (np.sin(x / 7.0) * np.cos(y / 11.0) * 0.5 + 0.5)[..., None] * np.array([255, 200, 150])
    img = base + rng.normal(0, 40, (h, w, 3))
    img[rng.random((h, w)) < 0.05] = 255
    img[rng.random((h, w)) < 0.05] = 0
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("src,dst", [((97, 131), (64, 64)), ((300, 200), (128, 192)), ((64, 48), (128, 96)),
                                     ((200, 100), (200, 64)), ((100, 200), (64, 200)), ((513, 767), (576, 832)),
                                     ((33, 47), (33, 47)), ((1920, 1080), (1024, 576)), ((301, 203), (77, 51)),
                                     ((4096, 64), (64, 64)), ((20000, 8), (64, 8))])
@pytest.mark.parametrize("kind", [R.LANCZOS, R.BILINEAR])
def test_resize_bit_exact_vs_pillow(ctx, src, dst, kind):
    rng = np.random.default_rng(src[0] * 31 + dst[0])
    img = rnd_img(rng, *src)
    got = ctx.resize_u8(torch.from_numpy(img).cuda(), dst, None, kind).cpu().numpy()
    want = np.asarray(Image.fromarray(img).resize(dst, Image.LANCZOS if kind == R.LANCZOS else Image.BILINEAR))
    assert np.array_equal(got, want), (np.abs(got.astype(int) - want.astype(int)).max(), (got != want).mean())
    assert np.array_equal(R.resize_u8(img, dst[0], dst[1], kind), want)


@pytest.mark.parametrize("src,bucket", [((640, 360), (576, 832)), ((300, 500), (768, 512)), ((512, 512), (512, 512)),
                                        ((401, 399), (1024, 1024)), ((3000, 2000), (1024, 704)),
                                        ((1000, 3000), (512, 1024))])
def test_smart_resize_matches_reference_transform(ctx, src, bucket):
    """Same result as the reference-shaped SmartResize on a PIL image (modules.py:142-178)."""
    rng = np.random.default_rng(src[0] + bucket[1])
    img = rnd_img(rng, *src)
    want = np.asarray(M.SmartResize(*bucket)(Image.fromarray(img)))
    got = gpu_smart_resize(torch.from_numpy(img).cuda(), *bucket, ctx=ctx).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(R.smart_resize_u8(img, *bucket), want)


def test_strided_source_and_batch_slot_output(ctx):
    rng = np.random.default_rng(1)
    wide = torch.from_numpy(rnd_img(rng, 700, 300)).cuda()
    view = wide[:, 100:600]                                   # row stride 2100 bytes, width 500
    batch = torch.zeros(3, 128, 192, 3, dtype=torch.uint8, device="cuda")
    ctx.resize_u8(view, (192, 128), (10, 20, 490, 280), R.LANCZOS, out=batch[1])
    want = np.asarray(Image.fromarray(view.cpu().numpy()).crop((10, 20, 490, 280)).resize((192, 128), Image.LANCZOS))
    assert np.array_equal(batch[1].cpu().numpy(), want)
    assert batch[0].abs().sum().item() == 0 and batch[2].abs().sum().item() == 0     # neighbours untouched
    with pytest.raises(_native.NativeError):
        ctx.resize_u8(view, (64, 64), (0, 0, 501, 300))        # box outside the image
    with pytest.raises(_native.NativeError):
        ctx.resize_u8(view, (64, 64), None, 7)                 # unknown filter


def test_batch_call_equals_single_calls(ctx):
    rng = np.random.default_rng(4)
    srcs = [torch.from_numpy(rnd_img(rng, w, h)).cuda() for w, h in ((640, 360), (500, 700), (300, 300))]
    boxes = [_native.smart_crop_box(s.shape[1], s.shape[0], 192, 128) for s in srcs]
    batch = ctx.resize_u8_batch(srcs, (192, 128), boxes)
    for i, (s, b) in enumerate(zip(srcs, boxes)):
        assert torch.equal(batch[i], ctx.resize_u8(s, (192, 128), b))
        want = np.asarray(M.SmartResize(192, 128)(Image.fromarray(s.cpu().numpy())))
        assert np.array_equal(batch[i].cpu().numpy(), want)


def test_properties_at_full_size(ctx):
    """Size-independent properties at the bench resolution: a constant image stays constant, identity size is
    a copy, and the result does not depend on how the image is batched."""
    const = torch.full((3000, 4000, 3), 137, dtype=torch.uint8, device="cuda")
    out = ctx.resize_u8(const, (1024, 1024))
    assert out.min().item() == 137 and out.max().item() == 137
    rng = np.random.default_rng(2)
    img = torch.from_numpy(rnd_img(rng, 1024, 1024)).cuda()
    assert torch.equal(ctx.resize_u8(img, (1024, 1024)), img)
    a = ctx.resize_u8(img, (704, 576))
    b = ctx.resize_u8(img, (704, 576))
    assert torch.equal(a, b)


def test_bucket_batcher_feeds_encoder_like_the_host_transform(ctx):
    """infer_full.py:97-98 on the host (PIL + ToTensor + Normalize) vs upload-uint8 + GPU resize + fused
    normalise: identical pixels, hence fp32-mode latents equal to rounding (rel L2 <= 1e-5)."""
    from vae_tagger_b200 import diffusers_vae_loader as L

    torch.manual_seed(0)
    wrap = L.DiffusersVAEWrapper(L.load_diffusers_vae_from_config(L.get_diffusers_vae_config())).cuda().eval()
    wrap.vae.precision = "fp32"
    rng = np.random.default_rng(3)
    sizes = [(640, 360), (360, 640), (500, 500), (800, 600), (600, 800), (1280, 720)]
    images = [rnd_img(rng, w, h) for w, h in sizes]
    arb = M.AspectRatioBucketing(128, 256, 64)                # small buckets: the fp32 verification mode is slow
    bb = BucketBatcher("cuda", batch_size=2, bucketing=arb)
    seen = {}
    for shape, keys, batch in bb.batches(enumerate(images)):
        assert batch.dtype == torch.uint8 and tuple(batch.shape[1:]) == (shape[1], shape[0], 3)
        lat = wrap.encode(batch)
        for k, l in zip(keys, lat):
            seen[k] = (shape, l.cpu())
    assert sorted(seen) == list(range(len(images)))
    for i, img in enumerate(images):
        shape = arb.bucket_for_size(img.shape[1], img.shape[0])
        assert seen[i][0] == shape
        tf = M.get_image_transform(0, True, shape)
        x = tf(Image.fromarray(img)).unsqueeze(0).cuda()
        want = wrap.encode(x)[0].cpu()
        rel = ((seen[i][1] - want).norm() / want.norm()).item()
        assert rel < 1e-5, (i, rel)
    # square path: transforms.Resize((res,res)) == BILINEAR
    sq = gpu_square_resize(torch.from_numpy(images[0]).cuda(), 192, ctx=ctx).cpu().numpy()
    from torchvision import transforms

    want = np.asarray(transforms.Resize((192, 192))(Image.fromarray(images[0])))
    assert np.array_equal(sq, want)


@pytest.mark.parametrize("src,dst", [((1, 1), (5, 7)), ((7, 5), (1, 1)), ((2, 300), (64, 3)), ((3, 3), (3, 3)),
                                     ((2, 300), (3, 150)), ((2, 201), (2, 100)), ((3, 300), (64, 3)), ((4, 2000), (64, 640))])
def test_degenerate_sizes(ctx, src, dst):
    """1-pixel images, and Pillow's vertical-pass-first rule for images more than 100 times taller than wide."""
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (src[1], src[0], 3), dtype=np.uint8)
    for kind, pil in ((R.LANCZOS, Image.LANCZOS), (R.BILINEAR, Image.BILINEAR)):
        got = ctx.resize_u8(torch.from_numpy(img).cuda(), dst, None, kind).cpu().numpy()
        assert np.array_equal(got, np.asarray(Image.fromarray(img).resize(dst, pil)))
