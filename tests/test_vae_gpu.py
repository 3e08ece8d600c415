"""AutoencoderKL (SD-1.5 VAE) on the seer_b200 kernels (-m gpu) against the CPU oracle oracle/vae_oracle.py on identical seeded
weights and inputs (`vae.decode` of the sampled latents, utils/ddim_sampling_utils.py:37-41, and `vae.encode`
of the reference frames, inference.py:186-187).  Tolerance: rel-L2 <= 2e-2 on decoded pixels / latent moments (the bf16
budget of north_star); the oracle itself is "parity unpinned" (diffusers is a third-party dependency absent from this image:
the anchors are the published checkpoint's state-dict schema and the reference's call sites)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import seer_oracle as so, vae_oracle as vo  # noqa: E402  (checkers only)
from seervideoldm_b200 import AutoencoderKL, ops  # noqa: E402
from seervideoldm_b200.pipeline import decode_latents  # noqa: E402
from seervideoldm_b200.vae import random_vae_state_dict  # noqa: E402

TOL = 2e-2
DEV = "cuda"


def gen(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.fixture(scope="module")
def model():
    sd = random_vae_state_dict(seed=0)
    vae = AutoencoderKL()
    vae.load_state_dict(sd, strict=True)
    return vae.cuda().eval(), sd


# ------------------------------------------------------------------------------------------------ kernels the VAE adds
def test_softmax_rows():
    s = gen(1, 300, 1024).to(DEV) * 30
    p = ops.softmax_rows(s, 0.044)
    ref = torch.softmax(s.double() * 0.044, -1).float()
    assert p.dtype == torch.bfloat16 and so.rel_l2(p.float().cpu(), ref.cpu()) < 4e-3
    assert torch.allclose(p.float().sum(-1), torch.ones(300, device=DEV), atol=2e-2)


@pytest.mark.parametrize("n_img,H,W,Cin,Cout", [(2, 8, 256, 64, 128), (1, 256, 256, 128, 128), (3, 4, 512, 64, 64)])
def test_conv3x3_wide_images(n_img, H, W, Cin, Cout):
    """Images wider than one 128-pixel tile (the VAE's 256x256 level): a tile is a 128-pixel row segment."""
    from seervideoldm_b200.packing import pack_conv3x3
    x = (gen(2, n_img, H, W, Cin)).to(DEV).bfloat16()
    w = gen(3, Cout, Cin, 3, 3) * (9 * Cin) ** -0.5
    b = gen(4, Cout).to(DEV)
    r = ops.gemm_ex(None, pack_conv3x3(w).to(DEV), x_img=x, bias=b, col_stats=True)
    assert r is not None
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(DEV).bfloat16().float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert so.rel_l2(r.out.cpu(), ref.cpu()) < 2e-5
    slabs = r.out.double().reshape(-1, 32, Cout)
    assert torch.allclose(r.col_stats[..., 0].double(), slabs.sum(1), rtol=1e-5, atol=1e-3)


def test_conv_out_wide_rows():
    from seervideoldm_b200.packing import pack_conv_out
    for (n, H, W, cin, cout) in [(2, 6, 256, 128, 3), (1, 5, 100, 64, 4), (1, 3, 130, 64, 2)]:
        h = gen(5, n * H * W, cin).to(DEV)
        wo, bo = gen(6, cout, cin, 3, 3) * 0.05, gen(7, cout).to(DEV)
        out = ops.conv_out(h, pack_conv_out(wo).to(DEV), bo, n, 1, H, W)
        ref = F.conv2d(h.double().reshape(n, H, W, cin).permute(0, 3, 1, 2), wo.to(DEV).double(), bo.double(), padding=1).float()
        assert out.shape == (n, cout, 1, H, W) and so.rel_l2(out[:, :, 0].cpu(), ref.cpu()) < 1e-5, (n, H, W, cin, cout)


def test_downsample_pad_right_bottom():
    """Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a stride-2 conv without padding == strided TMA taps at offset 0."""
    from seervideoldm_b200.packing import pack_conv3x3
    n, H, C, Cout = 3, 32, 128, 128
    x = gen(8, n, H, H, C).to(DEV).bfloat16()
    w, b = gen(9, Cout, C, 3, 3) * (9 * C) ** -0.5, gen(10, Cout).to(DEV)
    r = ops.gemm_ex(None, pack_conv3x3(w).to(DEV), x_img=x, conv_stride=2, conv_taps=(3, 3, 0, 0), bias=b)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.to(DEV).bfloat16().float(), b, stride=2)
    assert so.rel_l2(r.out.cpu(), ref.permute(0, 2, 3, 1).reshape(-1, Cout).cpu()) < 2e-5


# ------------------------------------------------------------------------------------------------ network parity
@pytest.mark.parametrize("n,h", [(2, 8), (3, 16)])
def test_decode_vs_oracle(model, n, h):
    vae, sd = model
    z = gen(20 + h, n, 4, h, h) * 3.0
    ref = vo.decode(sd, z)
    out = vae.decode(z.cuda()).sample
    assert out.shape == ref.shape and out.dtype == torch.float32
    err = so.rel_l2(out.cpu(), ref)
    print(f"VAE decode n={n} latent {h}x{h} -> {8 * h}x{8 * h}: rel-L2 {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("n,H", [(2, 64), (1, 128)])
def test_encode_vs_oracle(model, n, H):
    vae, sd = model
    x = gen(30 + H, n, 3, H, H)
    ref = vo.encode_moments(sd, x)
    dist = vae.encode(x.cuda()).latent_dist
    err = so.rel_l2(dist.parameters.cpu(), ref)
    print(f"VAE encode n={n} {H}x{H}: moments rel-L2 {err:.3e}")
    assert dist.parameters.shape == ref.shape and err < TOL
    noise_gen = torch.Generator(device="cuda").manual_seed(7)
    s = dist.sample(noise_gen)
    noise = torch.randn(dist.mean.shape, generator=torch.Generator(device="cuda").manual_seed(7), device="cuda")
    assert torch.allclose(s, dist.mean + dist.std * noise)
    assert so.rel_l2(dist.mode().cpu(), ref.chunk(2, 1)[0]) < TOL


@pytest.mark.slow
def test_decode_full_size_and_pipeline(model):
    """256x256 frames (32x32 latents: the benchmark's clip geometry) through `decode_latents`, the second half of the reference's
    ddim_sample (ddim_sampling_utils.py:37-41): exercises the 256-pixel-wide conv tiles, the d = 512 attention at L = 1024 and
    the upsample phase convs at every level."""
    vae, sd = model
    lat = gen(40, 1, 4, 2, 32, 32) * 0.18215 * 3.0            # (n c f h w) latents in the sampler's scale
    ref = vo.decode_latents(sd, lat)
    out = decode_latents(vae, lat.cuda())
    assert out.shape == (1, 3, 2, 256, 256)
    raw = vae.decode((1 / 0.18215 * lat.permute(0, 2, 1, 3, 4).reshape(2, 4, 32, 32)).cuda()).sample
    raw_ref = vo.decode(sd, 1 / 0.18215 * lat.permute(0, 2, 1, 3, 4).reshape(2, 4, 32, 32))
    err = so.rel_l2(raw.cpu(), raw_ref)
    print(f"VAE decode 2 frames 32x32 -> 256x256: rel-L2 {err:.3e}; clamped pixels max abs diff {float((out.cpu() - ref).abs().max()):.3e}")
    assert err < TOL
    assert float((out.cpu() - ref).abs().max()) < 0.1


@pytest.mark.slow
def test_encode_full_size(model):
    vae, sd = model
    x = gen(41, 1, 3, 256, 256)
    ref = vo.encode_moments(sd, x)
    err = so.rel_l2(vae.encode(x.cuda()).latent_dist.parameters.cpu(), ref)
    print(f"VAE encode 256x256: moments rel-L2 {err:.3e}")
    assert err < TOL
