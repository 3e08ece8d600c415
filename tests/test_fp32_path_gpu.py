"""fp32-parity path (-m gpu): north_star's "rel-L2 <= 1e-4 for the fp32 path".

Kernel level: the split-operand tcgen05 GEMM / conv against fp64 PyTorch on the SAME fp32 inputs (no pre-rounding: the
point of the path is that bf16 rounding of the operands no longer shows), and the fp32 SIMT kernels of csrc/fp32_path.cu.
Step level: SeerUNet.set_precision("fp32") against the oracle and against golden eps of the unmodified reference."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import seer_oracle as so  # noqa: E402  (checker only)
from seervideoldm_b200 import DDIMSampler, SeerUNet, ops, packing, unet_fp32  # noqa: E402
from seervideoldm_b200.config import sd15_config  # noqa: E402
from seervideoldm_b200.weights import random_state_dict  # noqa: E402

DEV = "cuda"
STEP_TOL_FP32 = 1e-4          # BASELINE.json north_star
KERNEL_TOL = 2e-5             # dropped lo*lo term + 16-bit operand mantissas: ~4e-6 expected


def rel(a, b):
    return so.rel_l2(a.detach().double().cpu(), b.detach().double().cpu())


def rn(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def test_split3_layout_and_exactness():
    x = rn(1, 96, 128, scale=3.0)
    s = ops.split3(x)
    assert s.shape == (96, 384) and s.dtype == torch.bfloat16
    hi = x.bfloat16()
    assert torch.equal(s[:, :128], hi) and torch.equal(s[:, 128:256], hi)
    assert torch.equal(s[:, 256:], (x - hi.float()).bfloat16())
    assert ((s[:, :128].float() + s[:, 256:].float()) - x).abs().max() <= x.abs().max() * 2.0 ** -16
    # concat placement: two parts into one [M, 3*(C1+C2)] buffer
    y = rn(2, 96, 64)
    buf = torch.zeros(96, 3 * 192, device=DEV, dtype=torch.bfloat16)
    ops.split3(x, out=buf, col0=0, ctot=192)
    ops.split3(y, out=buf, col0=128, ctot=192)
    ref = ops.split3(torch.cat([x, y], 1).contiguous())
    assert torch.equal(buf, ref)
    # fused nearest 2x upsample
    img = rn(3, 2 * 4 * 4, 64)
    up = ops.split3(img, upsample=(2, 4, 4)).view(2, 8, 8, 192)
    ref = ops.split3(img).view(2, 4, 4, 192).repeat_interleave(2, 1).repeat_interleave(2, 2)
    assert torch.equal(up, ref)


@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (1000, 640, 768), (4096, 2560, 320), (512, 320, 1280)])
def test_linear32(M, N, K):
    x, w = rn(10, M, K), rn(11, N, K, scale=K ** -0.5)
    bias, res = rn(12, N), rn(13, M, N)
    out = unet_fp32.linear32(x, unet_fp32._w3_linear(w), bias, residual=res)
    ref = x.double() @ w.double().t() + bias.double() + res.double()
    err = rel(out, ref)
    print(f"linear32 M={M} N={N} K={K}: rel-L2 {err:.2e}")
    assert out.dtype == torch.float32 and err < KERNEL_TOL
    # the plain bf16 GEMM on the same fp32 inputs sits three orders of magnitude higher
    plain = ops.gemm(x.bfloat16(), w.bfloat16(), bias=bias, residual=res)
    assert rel(plain, ref) > 20 * err


@pytest.mark.parametrize("n_img,H,Cin,Cout,Csc", [(3, 16, 64, 128, 0), (2, 32, 128, 64, 0), (4, 8, 64, 128, 192)])
def test_conv3x3_32(n_img, H, Cin, Cout, Csc):
    x = rn(20, n_img, H, H, Cin)
    w = rn(21, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    bias = rn(22, Cout)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=1)
    kw = {}
    wsc = None
    if Csc:
        raw = rn(23, n_img * H * H, Csc)
        wsc = rn(24, Cout, Csc, 1, 1, scale=Csc ** -0.5)
        ref = ref + F.conv2d(raw.view(n_img, H, H, Csc).permute(0, 3, 1, 2).double(), wsc.double())
        kw["a2"] = ops.split3(raw)
    out = unet_fp32.conv3x3_32(x.reshape(-1, Cin), n_img, H, H, unet_fp32._w3_conv3x3(w, wsc), bias=bias, **kw)
    err = rel(out, ref.permute(0, 2, 3, 1).reshape(-1, Cout))
    print(f"conv3x3_32: rel-L2 {err:.2e}")
    assert err < KERNEL_TOL


def test_layernorm_geglu_f32():
    x = rn(30, 333, 640, scale=2.0) + 0.5
    g, b = rn(31, 640), rn(32, 640)
    y = ops.layernorm_f32(x, g, b)
    assert rel(y, F.layer_norm(x.double(), (640,), g.double(), b.double(), 1e-5)) < 1e-6
    h = rn(33, 257, 2 * 1280, scale=1.5)
    out = ops.geglu_f32(h)
    ref = h[:, :1280].double() * F.gelu(h[:, 1280:].double())
    assert rel(out, ref) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rope_ex(dtype):
    heads, d, T, Fr = 8, 96, 77, 5                      # FSText geometry: position = frame index = (row // 77) % F
    C = heads * d
    M = 2 * Fr * T
    freqs = 1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32)).to(DEV)
    qkv = rn(40, M, 3 * C).to(dtype)
    ref = qkv.clone().float()
    pos = (torch.arange(M, device=DEV) // T) % Fr
    for col in (0, C):
        blk = ref[:, col:col + C].reshape(M, heads, d)
        blk[:, :, :32] = _rope_ref(blk[:, :, :32], pos, freqs)
        ref[:, col:col + C] = blk.reshape(M, C)
    ops.rope_ex(qkv, T, Fr, heads, d, 0, C, freqs)
    assert rel(qkv.float(), ref) < (1e-6 if dtype == torch.float32 else 4e-3)
    assert torch.equal(qkv[:, 2 * C:].float(), ref[:, 2 * C:])            # V untouched


def _rope_ref(x, pos, freqs):
    """Interleaved-pair rotation (rotary-embedding-torch 0.1.5 rotate_half): x [M, heads, 2*nf]."""
    ang = pos[:, None].double() * freqs[None, :].double()              # [M, nf]
    cs, sn = ang.cos()[:, None, :], ang.sin()[:, None, :]
    xe, xo = x[..., 0::2].double(), x[..., 1::2].double()
    out = torch.empty_like(x, dtype=torch.float64)
    out[..., 0::2] = xe * cs - xo * sn
    out[..., 1::2] = xo * cs + xe * sn
    return out.float()


def _attn_ref(q, k, v, causal):
    s = (q.double() @ k.double().transpose(-1, -2)) * q.shape[-1] ** -0.5
    if causal:
        L = s.shape[-1]
        s = s.masked_fill(~torch.ones(L, L, dtype=torch.bool, device=s.device).tril(), float("-inf"))
    return s.softmax(-1) @ v.double()


@pytest.mark.parametrize("d,frames,Lq,Lk", [(40, 2, 1024, 1024), (80, 3, 256, 256), (160, 2, 64, 64), (160, 2, 16, 16), (40, 2, 100, 77),
                                            (96, 4, 77, 77), (96, 1, 5 * 77, 77)])
def test_attention_f32_dense(d, frames, Lq, Lk):
    heads = 8
    C = heads * d
    q = rn(50, frames * Lq, C)
    kv = rn(51, frames * Lk, 2 * C)
    mode = ops.ATTN_SPATIAL if Lq == Lk else ops.ATTN_CROSS
    out = ops.attention(q, kv[:, :C], kv[:, C:], mode=mode, heads=heads, n_outer=frames, Lq=Lq, Lk=Lk)
    qh = q.reshape(frames, Lq, heads, d).permute(0, 2, 1, 3)
    kh = kv[:, :C].reshape(frames, Lk, heads, d).permute(0, 2, 1, 3)
    vh = kv[:, C:].reshape(frames, Lk, heads, d).permute(0, 2, 1, 3)
    ref = _attn_ref(qh, kh, vh, False).permute(0, 2, 1, 3).reshape(frames * Lq, C)
    assert out.dtype == torch.float32 and rel(out, ref) < 2e-6


@pytest.mark.parametrize("d,B,Fr,H", [(40, 1, 3, 32), (80, 2, 4, 16), (160, 2, 3, 8), (160, 1, 5, 4)])
def test_attention_f32_scta(d, B, Fr, H):
    heads = 8
    C = heads * d
    T = Fr * H * H
    qkv = rn(52, B * T, 3 * C)
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=Fr, H=H, W=H)
    t = qkv.reshape(B, T, 3, heads, d).permute(2, 0, 3, 1, 4)
    ref = torch.empty(B, heads, T, d, device=DEV, dtype=torch.float64)
    for seq in torch.from_numpy(so.scta_sequences(Fr, H, H)).to(DEV):
        ref[:, :, seq] = _attn_ref(t[0][:, :, seq], t[1][:, :, seq], t[2][:, :, seq], True)
    assert rel(out, ref.permute(0, 2, 1, 3).reshape(B * T, C)) < 2e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 6e-3)])
def test_attention_frame_mode(dtype, tol):
    """FSText temporal block: causal attention along the frame axis, one sequence per (clip, token) (attention.py:393-396)."""
    heads, d, B, Fr, T = 8, 96, 2, 6, 77
    C = heads * d
    qkv = rn(53, B * Fr * T, 3 * C).to(dtype)
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_FRAME, heads=heads, n_outer=B, F=Fr, H=T)
    t = qkv.float().reshape(B, Fr, T, 3, heads, d).permute(3, 0, 2, 4, 1, 5)          # (3, B, T, heads, Fr, d)
    ref = _attn_ref(t[0], t[1], t[2], True).permute(0, 3, 1, 2, 4).reshape(B * Fr * T, C)   # (B, Fr, T, heads, d)
    assert rel(out.float(), ref) < tol


# ------------------------------------------------------------------------------------------------ step level
@pytest.fixture(scope="module")
def model():
    cfg = sd15_config(sample_size=32)
    sd = random_state_dict(cfg, seed=0)
    net = SeerUNet(sample_size=32, cross_attention_dim=768)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval().set_precision("fp32")
    return net, sd


def gen(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("B,Fr,H,cf,tval", [(1, 2, 8, 0, 991), (2, 3, 16, 0, 496), (1, 4, 16, 2, 1), (2, 2, 32, 1, 750)])
def test_fp32_step_vs_oracle(model, B, Fr, H, cf, tval):
    net, sd = model
    x, c = gen(11, B, 4, Fr, H, H), gen(12, B, Fr, 77, 768)
    t = torch.full((B,), tval, dtype=torch.long)
    ref = so.unet_forward(sd, x, t, c, cf)
    out = net(x.cuda(), t.cuda(), c.cuda(), cond_frame=cf)
    err = so.rel_l2(out.cpu(), ref)
    print(f"fp32 step rel-L2 B={B} F={Fr} H={H} cond_frame={cf}: {err:.3e}")
    assert out.dtype == torch.float32 and err < STEP_TOL_FP32


def test_fp32_step_vs_reference_golden(model, golden_dir):
    """Golden eps produced by the UNMODIFIED reference in fp32 (oracle/make_golden.py)."""
    net, _ = model
    g = torch.load(os.path.join(golden_dir, "unet_sd15.pt"), weights_only=False)
    for case in g["cases"]:
        B, Fr, H = case["shape"]
        x, c = gen(case["x_seed"], B, 4, Fr, H, H), gen(case["c_seed"], B, Fr, 77, 768)
        out = net(x.cuda(), case["t"].cuda(), c.cuda(), cond_frame=case["cond_frame"])
        err = so.rel_l2(out.cpu(), case["y"])
        print(f"fp32 step vs reference golden {case['shape']}: {err:.3e}")
        assert err < STEP_TOL_FP32


def test_fp32_precision_switch_and_graph(model):
    """set_precision flips the kernels under the same module; the DDIM sampler's CUDA graph is keyed on it."""
    net, sd = model
    b, F2, H = 1, 2, 8
    xT, c = gen(51, b, 4, F2, H, H), gen(52, b, F2, 77, 768)
    uc = gen(53, b, F2, 77, 768)
    sampler = DDIMSampler(torch.device("cuda"))
    ref, _ = so.ddim_sample_latents(lambda x, t, cc, cf: so.unet_forward(sd, x, t, cc, cf), xT, c, None, 10, 7.5, uc)
    from seervideoldm_b200.pipeline import ddim_sample_latents
    lat32 = ddim_sample_latents(sampler, net, (b, 4, F2, H, H), c.cuda(), xT.cuda(), None, ddim_steps=10, scale=7.5, uc=uc.cuda())
    e32 = so.rel_l2(lat32.cpu(), ref)
    try:
        net.set_precision("bf16")
        lat16 = ddim_sample_latents(sampler, net, (b, 4, F2, H, H), c.cuda(), xT.cuda(), None, ddim_steps=10, scale=7.5, uc=uc.cuda())
    finally:
        net.set_precision("fp32")
    e16 = so.rel_l2(lat16.cpu(), ref)
    print(f"11-evaluation DDIM+CFG final latents: fp32 path {e32:.3e}, bf16 path {e16:.3e}")
    assert e32 < 1e-3 and e16 < 5e-2 and e32 < e16 / 10


@pytest.mark.slow
def test_fp32_full_size_sthv2_step(model):
    """BASELINE.json config 1/2 shape: UNet batch 2 (CFG), 12 frames, 32x32 latent, t = 991 — fp32 path."""
    net, sd = model
    x, c = gen(61, 2, 4, 12, 32, 32), gen(62, 2, 12, 77, 768)
    t = torch.full((2,), 991, dtype=torch.long)
    ref = so.unet_forward(sd, x, t, c, 0)
    out = net(x.cuda(), t.cuda(), c.cuda())
    err = so.rel_l2(out.cpu(), ref)
    print(f"fp32 full-size Sthv2 step rel-L2: {err:.3e}")
    assert err < STEP_TOL_FP32
