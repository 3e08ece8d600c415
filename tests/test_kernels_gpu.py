"""Kernel-level parity (-m gpu): every C-ABI entry point against a plain PyTorch fp32 statement of the same op
on the same seeded inputs.  bf16 operands are rounded once and the reference is computed in fp32/fp64 from the
ROUNDED values, so the only difference left is accumulation order (tolerances are written per test)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import seer_oracle as so  # noqa: E402  (checker only)
from seervideoldm_b200 import ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    return so.rel_l2(a.detach().cpu(), b.detach().cpu())


def rn(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 160, 64), (256, 320, 320), (384, 640, 1280), (192, 1280, 2560), (1000, 128, 192),
                                   (24576, 320, 320), (77 * 24, 640, 768), (128, 64, 64)])
def test_gemm_plain(M, N, K):
    a = rn(1, M, K).bfloat16()
    w = rn(2, N, K, scale=K ** -0.5).bfloat16()
    out = ops.gemm(a, w)
    ref = a.float() @ w.float().t()
    assert out.dtype == torch.float32
    assert rel(out, ref) < 2e-5


def test_gemm_epilogues():
    M, N, K, K2 = 512, 320, 640, 320
    a, a2 = rn(3, M, K).bfloat16(), rn(4, M, K2).bfloat16()
    w = rn(5, N, K + K2, scale=(K + K2) ** -0.5).bfloat16()
    bias, res = rn(6, N), rn(7, M, N)
    ref = torch.cat([a, a2], 1).float() @ w.float().t() + bias + res
    out = ops.gemm(a, w, a2=a2, bias=bias, residual=res)
    assert rel(out, ref) < 2e-5
    out16 = ops.gemm(a, w, a2=a2, bias=bias, residual=res, out_dtype=torch.bfloat16)
    assert out16.dtype == torch.bfloat16 and rel(out16.float(), ref) < 4e-3
    # per-sample bias rows (time-embedding add): 4 samples of 128 rows, bias matrix is a column slice (ldb > N)
    wide = rn(8, 4, 2 * N)
    pb = wide[:, N:]
    out = ops.gemm(a, w[:, :K].contiguous(), bias=pb, bias_div=128)
    ref = a.float() @ w[:, :K].float().t() + pb.repeat_interleave(128, 0)
    assert rel(out, ref) < 2e-5
    # A given as a column slice of a wider buffer (lda > K), out written into a column slice
    wide_a = rn(9, M, 3 * K).bfloat16()
    dst = torch.zeros(M, 2 * N, device=DEV)
    ops.gemm(wide_a[:, K:2 * K], w[:, :K].contiguous(), out=dst[:, N:])
    assert rel(dst[:, N:], wide_a[:, K:2 * K].float() @ w[:, :K].float().t()) < 2e-5
    assert float(dst[:, :N].abs().max()) == 0.0


def test_gemm_geglu():
    M, C = 384, 320
    a = rn(10, M, C).bfloat16()
    w = rn(11, 8 * C, C, scale=C ** -0.5)
    b = rn(12, 8 * C)
    u = a.float() @ w.bfloat16().float().t() + b
    ref = u[:, :4 * C] * F.gelu(u[:, 4 * C:])
    from seervideoldm_b200.packing import pack_geglu
    wp, bp = pack_geglu(w, b)
    out = ops.gemm(a, wp.to(DEV), bias=bp.to(DEV), geglu=True)
    assert out.shape == (M, 4 * C) and out.dtype == torch.bfloat16
    assert rel(out.float(), ref) < 4e-3


@pytest.mark.parametrize("M,N,K", [(20000, 640, 320), (4096 + 40, 960, 128), (2500, 1920, 192), (33000, 256, 64),
                                   (9000, 1280, 640), (640, 3840, 128)])
def test_gemm_persistent_many_tiles(M, N, K):
    """More output tiles than SMs (every CTA loops), ragged last M tile, every tile width the planner can pick."""
    a = rn(13, M, K).bfloat16()
    w = rn(14, N, K, scale=K ** -0.5).bfloat16()
    bias, res = rn(15, N), rn(16, M, N)
    ref = a.float() @ w.float().t() + bias + res
    out = ops.gemm(a, w, bias=bias, residual=res)
    assert rel(out, ref) < 2e-5
    out16 = ops.gemm(a, w, bias=bias, out_dtype=torch.bfloat16)
    assert rel(out16.float(), ref - res) < 4e-3


def test_gemm_v2_residual_dtypes_and_dual_outputs():
    M, N, K = 3000, 640, 320
    a = rn(17, M, K).bfloat16()
    w = rn(18, N, K, scale=K ** -0.5).bfloat16()
    bias = rn(19, N)
    res32, res16 = rn(20, M, N), rn(20, M, N).bfloat16()
    base = a.float() @ w.float().t() + bias
    # fp32 residual -> fp32 out + bf16 copy
    r = ops.gemm_ex(a, w, bias=bias, residual=res32, also_bf16=True)
    assert rel(r.out, base + res32) < 2e-5
    assert torch.equal(r.out16, r.out.bfloat16())
    # fp32 residual -> bf16 only
    r = ops.gemm_ex(a, w, bias=bias, residual=res32, out_dtype=torch.bfloat16)
    assert rel(r.out.float(), base + res32) < 4e-3
    # bf16 residual -> bf16 only (in place in the slot)
    r = ops.gemm_ex(a, w, bias=bias, residual=res16, out_dtype=torch.bfloat16)
    assert rel(r.out.float(), base + res16.float()) < 4e-3
    # bf16 residual -> fp32 (+ bf16 copy)
    r = ops.gemm_ex(a, w, bias=bias, residual=res16, also_bf16=True)
    assert rel(r.out, base + res16.float()) < 2e-5
    assert torch.equal(r.out16, r.out.bfloat16())
    # no residual, both outputs
    r = ops.gemm_ex(a, w, bias=bias, also_bf16=True)
    assert rel(r.out, base) < 2e-5 and torch.equal(r.out16, r.out.bfloat16())


@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (1000, 640, 128), (40, 320, 64)])
def test_gemm_v2_statistics_outputs(M, N, K):
    a = rn(21, M, K).bfloat16()
    w = rn(22, N, K, scale=K ** -0.5).bfloat16()
    bias, res = rn(23, N), rn(24, M, N)
    r = ops.gemm_ex(a, w, bias=bias, residual=res, col_stats=True, row_stats=True)
    v = r.out.double()
    # per-row partial sums add up to the row's (sum, sumsq)
    rs = r.row_stats.double().sum(0)
    assert torch.allclose(rs[:, 0], v.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(rs[:, 1], (v * v).sum(1), rtol=1e-5, atol=1e-3)
    # per-(32-row slab, column) sums
    S = (M + 31) // 32
    pad = torch.zeros(S * 32, N, dtype=torch.double, device=DEV)
    pad[:M] = v
    slabs = pad.reshape(S, 32, N)
    assert r.col_stats.shape == (S, N, 2)
    assert torch.allclose(r.col_stats[..., 0].double(), slabs.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(r.col_stats[..., 1].double(), (slabs * slabs).sum(1), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("geglu", [False, True])
def test_gemm_v2_layernorm_fold(geglu):
    """LayerNorm folded into the consuming GEMM: producer emits per-row (sum, sumsq); consumer runs on the RAW bf16
    stream with gamma folded into the weights and applies rstd * (acc - mean * colsum) + beta @ W.T in its epilogue."""
    M, C, N = 2000, 320, (1280 if geglu else 960)
    a0 = rn(25, M, C).bfloat16()
    w0 = (torch.eye(C, device=DEV) * 1.5 + rn(26, C, C, scale=0.02)).bfloat16()
    prod = ops.gemm_ex(a0, w0, bias=rn(27, C) * 0.3 + 0.2, also_bf16=True, row_stats=True)
    x, x16 = prod.out, prod.out16
    gamma, beta = rn(28, C) * 0.2 + 1.0, rn(29, C) * 0.1
    W, b = rn(30, N, C, scale=C ** -0.5), rn(31, N)
    wf = (W * gamma[None, :]).bfloat16()
    colsum = wf.float().sum(1).contiguous()
    bias_f = (W @ beta + b).contiguous()
    r = ops.gemm_ex(x16, wf.contiguous(), bias=bias_f, out_dtype=torch.bfloat16, geglu=geglu, ln=(prod.row_stats, colsum, 1e-5))
    u = F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5) @ W.double().t() + b.double()
    if geglu:
        # un-permuted reference: the test feeds weights whose rows are already in value/gate-32 block order
        u = u.reshape(M, N // 64, 2, 32)
        ref = (u[:, :, 0] * F.gelu(u[:, :, 1])).reshape(M, N // 2)
    else:
        ref = u
    assert rel(r.out.float(), ref.float()) < 8e-3


def test_groupnorm_from_gemm_statistics():
    """GroupNorm statistics taken from the producing GEMMs' epilogues (virtual concat of two producers)."""
    B, T, C1, C2, K = 2, 512, 640, 320, 128
    M = B * T
    p1 = ops.gemm_ex(rn(32, M, K).bfloat16(), rn(33, C1, K, scale=K ** -0.5).bfloat16(), bias=rn(34, C1), col_stats=True)
    p2 = ops.gemm_ex(rn(35, M, K).bfloat16(), rn(36, C2, K, scale=K ** -0.5).bfloat16(), bias=rn(37, C2) + 0.5, col_stats=True)
    g, b = rn(38, C1 + C2), rn(39, C1 + C2)
    y_ref = ops.groupnorm(p1.out, p2.out, B, g, b, 1e-5, True, out_dtype=torch.float32)
    y = ops.groupnorm(p1.out, p2.out, B, g, b, 1e-5, True, out_dtype=torch.float32, stats1=p1.col_stats, stats2=p2.col_stats)
    assert rel(y, y_ref) < 2e-6
    y1 = ops.groupnorm(p1.out, None, B, g[:C1].contiguous(), b[:C1].contiguous(), 1e-6, False, stats1=p1.col_stats)
    y1_ref = ops.groupnorm(p1.out, None, B, g[:C1].contiguous(), b[:C1].contiguous(), 1e-6, False)
    assert rel(y1.float(), y1_ref.float()) < 1e-3


def test_conv3x3_statistics_and_dual():
    n_img, H, Cin, Cout = 6, 16, 64, 320
    x = rn(40, n_img, H, H, Cin).bfloat16()
    w = rn(41, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    from seervideoldm_b200.packing import pack_conv3x3
    res = rn(42, n_img * H * H, Cout)
    r = ops.conv3x3_ex(x, pack_conv3x3(w).to(DEV), residual=res, col_stats=True, also_bf16=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout) + res
    assert rel(r.out, ref) < 2e-5
    assert torch.equal(r.out16, r.out.bfloat16())
    slabs = r.out.double().reshape(-1, 32, Cout)
    assert torch.allclose(r.col_stats[..., 0].double(), slabs.sum(1), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (3000, 640, 640), (1000, 1280, 1280), (262144 // 8, 320, 1280)])
def test_gemm_bf16_token_stream_kinds(M, N, K):
    """The epilogue kinds of the bf16 token stream inside a transformer block: proj_in (bf16 out + LayerNorm row sums),
    to_out (bf16 residual in, bf16 out, row sums), FF out (bf16 residual, bf16 out).  Row sums are taken from the fp32
    values before rounding."""
    a = rn(43, M, K).bfloat16()
    w = rn(44, N, K, scale=K ** -0.5).bfloat16()
    bias, res16 = rn(45, N), rn(46, M, N).bfloat16()
    base = a.double() @ w.double().t() + bias.double()
    r = ops.gemm_ex(a, w, bias=bias, out_dtype=torch.bfloat16, row_stats=True)
    assert r.out.dtype == torch.bfloat16 and rel(r.out.float(), base.float()) < 4e-3
    rs = r.row_stats.double().sum(0)
    assert torch.allclose(rs[:, 0], base.sum(1), rtol=1e-4, atol=2e-3)
    assert torch.allclose(rs[:, 1], (base * base).sum(1), rtol=1e-4, atol=2e-3)
    full = base + res16.double()
    r = ops.gemm_ex(a, w, bias=bias, residual=res16, out_dtype=torch.bfloat16, row_stats=True)
    assert rel(r.out.float(), full.float()) < 4e-3
    rs = r.row_stats.double().sum(0)
    assert torch.allclose(rs[:, 0], full.sum(1), rtol=1e-4, atol=2e-3)
    assert torch.allclose(rs[:, 1], (full * full).sum(1), rtol=1e-4, atol=2e-3)
    r2 = ops.gemm_ex(a, w, bias=bias, residual=res16, out_dtype=torch.bfloat16)
    assert torch.equal(r2.out, r.out)            # same values with and without the statistics output


@pytest.mark.parametrize("M,N,K", [(4096, 320, 1280), (2048, 640, 640), (262144 // 8, 320, 320)])
def test_gemm_bf16_residual_stream_kind(M, N, K):
    """proj_out / conv2 of the bf16 residual stream: bf16 residual in, bf16 out, GroupNorm column sums of the fp32 values —
    the same values and the same statistics as the fp32-output kind on the same operands."""
    a = rn(143, M, K).bfloat16()
    w = rn(144, N, K, scale=K ** -0.5).bfloat16()
    bias, res16 = rn(145, N), rn(146, M, N).bfloat16()
    r32 = ops.gemm_ex(a, w, bias=bias, residual=res16, col_stats=True)
    r16 = ops.gemm_ex(a, w, bias=bias, residual=res16, col_stats=True, out_dtype=torch.bfloat16)
    assert f"spec={128 | 16 | 32} " in ops.last_gemm_kernel()            # EK_POUT16: a compiled specialisation, not the generic path
    assert r16.out.dtype == torch.bfloat16 and torch.equal(r16.out, r32.out.bfloat16())
    full = a.double() @ w.double().t() + bias.double() + res16.double()
    assert rel(r16.out.float(), full.float()) < 4e-3
    # the column sums of a bf16-only output are those of the stored tensor (what the consuming GroupNorm normalises)
    slabs = r16.out.double().reshape(-1, 32, N)
    assert torch.allclose(r16.col_stats[..., 0].double(), slabs.sum(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(r16.col_stats[..., 1].double(), (slabs * slabs).sum(1), rtol=1e-5, atol=1e-4)
    assert rel(r16.col_stats, r32.col_stats) < 2e-3


def test_groupnorm_bf16_concat_sources_and_raw_copy():
    """GroupNorm over the virtual concat of two bf16 block outputs (bf16 residual stream) with producer statistics: equals
    GroupNorm of the fp32 tensors up to the bf16 rounding of the inputs; the raw copy is the bf16 concat itself."""
    B, T, C1, C2, K = 2, 512, 640, 320, 128
    M = B * T
    mk = lambda i, C, dt: ops.gemm_ex(rn(i, M, K).bfloat16(), rn(i + 1, C, K, scale=K ** -0.5).bfloat16(), bias=rn(i + 2, C) + 0.25,
                                      col_stats=True, out_dtype=dt)
    p1, p2 = mk(150, C1, torch.bfloat16), mk(153, C2, torch.bfloat16)
    q1, q2 = mk(150, C1, torch.float32), mk(153, C2, torch.float32)
    g, b = rn(156, C1 + C2), rn(157, C1 + C2)
    y16, raw = ops.groupnorm(p1.out, p2.out, B, g, b, 1e-5, True, want_raw=True, stats1=p1.col_stats, stats2=p2.col_stats)
    y32 = ops.groupnorm(q1.out, q2.out, B, g, b, 1e-5, True, stats1=q1.col_stats, stats2=q2.col_stats)
    assert y16.dtype == torch.bfloat16 and rel(y16.float(), y32.float()) < 6e-3
    assert torch.equal(raw, torch.cat([p1.out, p2.out], 1))
    with pytest.raises(ValueError, match="one dtype"):
        ops.groupnorm(p1.out, q2.out, B, g, b, 1e-5, True, stats1=p1.col_stats, stats2=q2.col_stats)


def test_conv_in_bf16_output():
    x = rn(160, 2, 4, 3, 16, 16)
    w, b = rn(161, 320, 36, scale=1 / 6.0), rn(162, 320)
    o32, s32 = ops.conv_in(x, w, b, col_stats=True)
    o16, s16 = ops.conv_in(x, w, b, col_stats=True, out_dtype=torch.bfloat16)
    assert o16.dtype == torch.bfloat16 and torch.equal(o16, o32.bfloat16()) and torch.equal(s16, s32)


def test_conv3x3_bf16_out_with_statistics():
    """conv1 of a ResNet block: bf16 output (read only by GroupNorm 2), column sums from the fp32 accumulators; the
    GroupNorm that consumes the bf16 tensor through those statistics equals GroupNorm of the fp32 conv output."""
    B, Fr, H, Cin, Cout = 2, 3, 16, 128, 320
    n_img = B * Fr
    x = rn(47, n_img, H, H, Cin).bfloat16()
    w = rn(48, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    temb = rn(49, B, Cout)
    from seervideoldm_b200.packing import pack_conv3x3
    wp = pack_conv3x3(w).to(DEV)
    T = Fr * H * H
    r32 = ops.conv3x3_ex(x, wp, bias=temb, bias_div=T, col_stats=True)
    r16 = ops.conv3x3_ex(x, wp, bias=temb, bias_div=T, col_stats=True, out_dtype=torch.bfloat16)
    assert r16.out.dtype == torch.bfloat16 and torch.equal(r16.out, r32.out.bfloat16())
    slabs = r16.out.double().reshape(-1, 32, Cout)          # column sums of the stored (bf16) tensor
    assert torch.allclose(r16.col_stats[..., 0].double(), slabs.sum(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(r16.col_stats[..., 1].double(), (slabs * slabs).sum(1), rtol=1e-5, atol=1e-4)
    assert rel(r16.col_stats, r32.col_stats) < 2e-3
    g, b = rn(50, Cout), rn(51, Cout)
    y32 = ops.groupnorm(r32.out, None, B, g, b, 1e-5, True, stats1=r32.col_stats)
    y16 = ops.groupnorm(r16.out, None, B, g, b, 1e-5, True, stats1=r16.col_stats)
    assert rel(y16.float(), y32.float()) < 6e-3
    ref = F.silu(F.group_norm(r32.out.reshape(B, T, Cout).permute(0, 2, 1).double(), 32, g.double(), b.double(), 1e-5))
    assert rel(y16.float(), ref.permute(0, 2, 1).reshape(B * T, Cout).float()) < 8e-3


@pytest.mark.parametrize("n_img,H,C,Cout", [(3, 16, 64, 160), (4, 32, 320, 320), (32, 8, 192, 320), (6, 16, 640, 640), (256, 8, 128, 160)])
def test_conv3x3_stride2_implicit(n_img, H, C, Cout):
    """Downsample3D (resnet.py:95-104): stride 2, pad 1, read straight from the full-resolution bf16 image through a TMA
    box with traversal stride 2 — against F.conv2d and against the explicit im2col path."""
    from seervideoldm_b200.packing import pack_conv3x3
    x = rn(52, n_img, H, H, C).bfloat16()
    w = rn(53, Cout, C, 3, 3, scale=(9 * C) ** -0.5)
    b = rn(54, Cout)
    wp = pack_conv3x3(w).to(DEV)
    r = ops.gemm_ex(None, wp, x_img=x, conv_stride=2, bias=b, col_stats=True)
    assert r is not None and r.out.shape == (n_img * (H // 2) ** 2, Cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel(r.out, ref) < 2e-5
    cols = ops.im2col3x3(x, stride=2)
    assert rel(r.out, ops.gemm(cols, wp, bias=b)) < 2e-6
    slabs = r.out.double().reshape(-1, 32, Cout)
    assert torch.allclose(r.col_stats[..., 0].double(), slabs.sum(1), rtol=1e-5, atol=1e-3)
    # the bf16 residual stream's form of the same launch: bf16 rows, column sums of the stored values
    r16 = ops.gemm_ex(None, wp, x_img=x, conv_stride=2, bias=b, col_stats=True, out_dtype=torch.bfloat16)
    assert torch.equal(r16.out, r.out.bfloat16())
    slabs16 = r16.out.double().reshape(-1, 32, Cout)
    assert torch.allclose(r16.col_stats[..., 0].double(), slabs16.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(r16.col_stats[..., 1].double(), (slabs16 * slabs16).sum(1), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("n_img,H,C,Cout", [(4, 8, 64, 160), (6, 16, 128, 320), (32, 4, 192, 160), (5, 16, 640, 640), (16, 8, 1280, 1280)])
def test_upsample_conv_phases(n_img, H, C, Cout):
    """Upsample3D (resnet.py:47-61): nearest 2x + conv3x3 as four 2x2-tap convs on the low-res image whose epilogues
    scatter the rows to the four pixel phases — against F.conv2d(F.interpolate(x)) with the ORIGINAL 3x3 weights."""
    from seervideoldm_b200.packing import pack_upsample_phases
    x = rn(55, n_img, H, H, C).bfloat16()
    w = rn(56, Cout, C, 3, 3, scale=(9 * C) ** -0.5)
    b = rn(57, Cout)
    phases = [p.to(DEV) for p in pack_upsample_phases(w.cpu())]
    M = n_img * 4 * H * H
    out = torch.full((M, Cout), float("nan"), device=DEV)
    st = torch.full((M // 32, Cout, 2), float("nan"), device=DEV)
    for ph in range(4):
        r = ops.gemm_ex(None, phases[ph], x_img=x, conv_taps=(2, 2, (ph & 1) - 1, (ph >> 1) - 1), up_phase=1 + ph, bias=b, out=out,
                        col_stats=st)
        assert r is not None
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert torch.isfinite(out).all() and torch.isfinite(st).all()
    # the phase weights are sums of 3x3 taps rounded to bf16 ONCE (the reference rounds each tap): bf16-weight-level agreement
    assert rel(out, ref.float()) < 4e-3
    # ... and exact (accumulation order only) against the same rounded phase weights applied as a dense fp32 conv
    outp = torch.empty(n_img, 2 * H, 2 * H, Cout, device=DEV, dtype=torch.double)
    xp = F.pad(x.double().permute(0, 3, 1, 2), (1, 1, 1, 1))
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        wk = phases[ph].double().reshape(Cout, C // 64, 2, 2, 64).permute(0, 1, 4, 2, 3).reshape(Cout, C, 2, 2)
        o = F.conv2d(xp[:, :, py:py + H + 1, px:px + H + 1], wk, b.double())
        outp[:, py::2, px::2] = o.permute(0, 2, 3, 1)
    assert rel(out, outp.reshape(-1, Cout).float()) < 2e-5
    # statistics: slab s of the output holds 32 consecutive OUTPUT-tensor rows?  No — slab 4*(m/32)+phase holds the phase rows of
    # 32 consecutive low-res pixels; what GroupNorm needs is that each sample's slabs sum to the sample's column sums
    grp = max(1, 32 // (H * H))              # a 32-pixel slab spans `grp` whole images when an image has fewer than 32 pixels
    S = st.double().reshape(n_img // grp, -1, Cout, 2).sum(1)
    v = out.double().reshape(n_img // grp, -1, Cout)
    assert torch.allclose(S[..., 0], v.sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(S[..., 1], (v * v).sum(1), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("n_img,H,C,Cout", [(6, 16, 128, 320), (16, 8, 1280, 1280)])
def test_upsample_conv_phases_bf16_stream(n_img, H, C, Cout):
    """The same four phase convs writing the bf16 residual stream: bf16 rows scattered to the pixel phases, column sums of the
    stored values in the phase-interleaved slab order."""
    from seervideoldm_b200.packing import pack_upsample_phases
    x = rn(155, n_img, H, H, C).bfloat16()
    w = rn(156, Cout, C, 3, 3, scale=(9 * C) ** -0.5)
    b = rn(157, Cout)
    phases = [p.to(DEV) for p in pack_upsample_phases(w.cpu())]
    M = n_img * 4 * H * H
    out32 = torch.empty((M, Cout), device=DEV)
    st32 = torch.empty((M // 32, Cout, 2), device=DEV)
    out16 = torch.full((M, Cout), float("nan"), device=DEV).bfloat16()
    st16 = torch.full((M // 32, Cout, 2), float("nan"), device=DEV)
    for ph in range(4):
        kw = dict(x_img=x, conv_taps=(2, 2, (ph & 1) - 1, (ph >> 1) - 1), up_phase=1 + ph, bias=b)
        assert ops.gemm_ex(None, phases[ph], out=out32, col_stats=st32, **kw) is not None
        assert ops.gemm_ex(None, phases[ph], out=out16, col_stats=st16, **kw) is not None
    assert torch.equal(out16, out32.bfloat16())
    assert torch.isfinite(st16).all() and rel(st16, st32) < 2e-3
    S = st16.double().reshape(n_img, -1, Cout, 2).sum(1)
    v = out16.double().reshape(n_img, -1, Cout)
    assert torch.allclose(S[..., 0], v.sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(S[..., 1], (v * v).sum(1), rtol=1e-5, atol=1e-2)


def test_gemm_320_wide_pair_tiles():
    """BN = 320 plan (one single-buffered 256 x 320 pair accumulator, two N = 160 UMMAs per A stage) — forced through the
    tuning hook on small shapes, and as the planner picks it at the level-0 / level-1 shapes of the benchmark."""
    from seervideoldm_b200.packing import pack_conv3x3
    try:
        ops.set_tuning("SEER_GEMM_BN", 320)
        for (M, N, K) in [(3000, 320, 1280), (2048 + 40, 640, 2560), (512, 640, 128)]:
            a = rn(58, M, K).bfloat16()
            w = rn(59, N, K, scale=K ** -0.5).bfloat16()
            bias, res = rn(60, N), rn(61, M, N)
            r = ops.gemm_ex(a, w, bias=bias, residual=res, col_stats=True, row_stats=True, also_bf16=True)
            assert "gemm_tc_kernel<320,2>" in ops.last_gemm_kernel(), ops.last_gemm_kernel()
            ref = a.float() @ w.float().t() + bias + res
            assert rel(r.out, ref) < 2e-5 and torch.equal(r.out16, r.out.bfloat16())
            v = r.out.double()
            assert torch.allclose(r.row_stats.double().sum(0)[:, 0], v.sum(1), rtol=1e-5, atol=1e-3)
            S = (M + 31) // 32
            pad = torch.zeros(S * 32, N, dtype=torch.double, device=DEV)
            pad[:M] = v
            assert torch.allclose(r.col_stats[..., 0].double(), pad.reshape(S, 32, N).sum(1), rtol=1e-5, atol=1e-3)
            r16 = ops.gemm_ex(a, w, bias=bias, residual=res.bfloat16(), out_dtype=torch.bfloat16)
            assert rel(r16.out.float(), ref - res + res.bfloat16().float()) < 4e-3
        n_img, H, Cin, Cout, Csc = 12, 16, 128, 320, 192
        x = rn(62, n_img, H, H, Cin).bfloat16()
        raw = rn(63, n_img * H * H, Csc).bfloat16()
        w, wsc = rn(64, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5), rn(65, Cout, Csc, 1, 1, scale=Csc ** -0.5)
        b = rn(66, Cout)
        r = ops.conv3x3_ex(x, pack_conv3x3(w, wsc).to(DEV), a2=raw, bias=b, col_stats=True)
        assert "gemm_tc_kernel<320,2>" in ops.last_gemm_kernel()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
        ref = ref + raw.float() @ wsc.reshape(Cout, Csc).bfloat16().float().t()
        assert rel(r.out, ref) < 2e-5
    finally:
        ops.set_tuning("SEER_GEMM_BN", None)
    # the planner's own choice at a benchmark shape (M = 65536 rows is enough for the persistent loop to wrap many times)
    M, N, K = 65536, 320, 1280
    a = rn(67, M, K).bfloat16()
    w = rn(68, N, K, scale=K ** -0.5).bfloat16()
    r = ops.gemm_ex(a, w, bias=rn(69, N), out_dtype=torch.bfloat16)
    print("planner at M=65536 N=320 K=1280:", ops.last_gemm_kernel())
    assert rel(r.out.float(), a.float() @ w.float().t() + rn(69, N)) < 4e-3


@pytest.mark.parametrize("d,Fr,hw,B", [(40, 4, 64, 2), (80, 2, 256, 1), (160, 4, 32, 3), (40, 16, 1024, 2)])
def test_gemm_qkv_rope_fused(d, Fr, hw, B):
    """RoPE (attention.py:649-651) fused into the LN-folded q/k/v projection's epilogue == the unfused projection followed by
    the stand-alone rotary pass, up to one bf16 rounding (the fused path rotates the fp32 accumulators)."""
    heads = 8
    C = heads * d
    T = Fr * hw
    M = B * T
    prod = ops.gemm_ex(rn(70, M, C).bfloat16(), (torch.eye(C, device=DEV) + rn(71, C, C, scale=0.02)).bfloat16(), bias=rn(72, C) * 0.1,
                       out_dtype=torch.bfloat16, row_stats=True)
    gamma, beta = rn(73, C) * 0.2 + 1.0, rn(74, C) * 0.1
    W = rn(75, 3 * C, C, scale=C ** -0.5)
    wf = (W * gamma[None, :]).bfloat16().contiguous()
    colsum, bias_f = wf.float().sum(1).contiguous(), (W @ beta).contiguous()
    freqs = (1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
    tab = ops.rope_table(freqs, T)
    ang = torch.arange(T, device=DEV).float()[:, None] * freqs[None, :]
    assert tab.dtype == torch.float16
    assert torch.allclose(tab[..., 0].float(), ang.cos(), atol=6e-4) and torch.allclose(tab[..., 1].float(), ang.sin(), atol=6e-4)
    fused = ops.gemm_ex(prod.out, wf, bias=bias_f, out_dtype=torch.bfloat16, ln=(prod.row_stats, colsum, 1e-5), rope=(tab, 2 * C, d)).out
    assert "spec=273" in ops.last_gemm_kernel(), ops.last_gemm_kernel()        # EK_QKV_ROPE = LN | OUT16 | ROPE
    # reference: the same projection in fp32, rotated in fp64 with the oracle's rotary statement
    plain32 = ops.gemm_ex(prod.out, wf, bias=bias_f, ln=(prod.row_stats, colsum, 1e-5)).out
    ref = plain32.clone().double().reshape(B, T, 3, heads, d)
    pos = torch.arange(T)
    for which in (0, 1):
        r = so.rope_interleaved(ref[:, :, which].permute(0, 2, 1, 3).cpu(), pos, 32)
        ref[:, :, which] = r.permute(0, 2, 1, 3).to(DEV)
    ref = ref.reshape(M, 3 * C).float()
    assert rel(fused.float(), ref) < 4e-3
    assert torch.equal(fused[:, 2 * C:], plain32[:, 2 * C:].bfloat16())                   # V untouched
    unfused = ops.gemm_ex(prod.out, wf, bias=bias_f, out_dtype=torch.bfloat16, ln=(prod.row_stats, colsum, 1e-5)).out
    ops.rope_inplace(unfused, T, heads, d, 0, C, freqs)
    assert rel(fused.float(), unfused.float()) < 6e-3


def test_gemm_rejects_bad_shapes():
    a = rn(1, 128, 100).bfloat16()
    w = rn(2, 160, 100).bfloat16()
    with pytest.raises(ValueError):
        ops.gemm(a, w)                      # K % 64 != 0
    with pytest.raises(TypeError):
        ops.gemm(a.float(), w)


# ------------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("n_img,H,Cin,Cout", [(3, 32, 64, 160), (2, 16, 320, 640), (5, 8, 128, 320), (11, 4, 192, 160),
                                             (24, 32, 320, 320), (3, 2, 64, 160), (130, 1, 64, 160), (2, 64, 64, 160)])
def test_conv3x3(n_img, H, Cin, Cout):
    x = rn(20, n_img, H, H, Cin).bfloat16()
    w = rn(21, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    b = rn(22, Cout)
    from seervideoldm_b200.packing import pack_conv3x3
    out = ops.conv3x3(x, pack_conv3x3(w).to(DEV), bias=b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel(out, ref) < 2e-5


def test_conv3x3_fused_shortcut_and_temb():
    n_img, H, Cin, Cout, Csc, frames = 4, 16, 128, 320, 192, 2
    x = rn(23, n_img, H, H, Cin).bfloat16()
    raw = rn(24, n_img * H * H, Csc).bfloat16()
    w, wsc = rn(25, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5), rn(26, Cout, Csc, 1, 1, scale=Csc ** -0.5)
    temb = rn(27, n_img // frames, Cout)
    from seervideoldm_b200.packing import pack_conv3x3
    wt = torch.cat([pack_conv3x3(w), wsc.reshape(Cout, Csc).bfloat16()], 1).contiguous().to(DEV)
    out = ops.conv3x3(x, wt, a2=raw, bias=temb, bias_div=frames * H * H)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    ref = ref + raw.float() @ wsc.reshape(Cout, Csc).bfloat16().float().t() + temb.repeat_interleave(frames * H * H, 0)
    assert rel(out, ref) < 2e-5


def test_conv3x3_generic_fallback_and_stride2():
    from seervideoldm_b200.packing import pack_conv3x3
    x = rn(28, 2, 12, 12, 64).bfloat16()          # 12 does not divide 128 -> im2col path
    w = rn(29, 160, 64, 3, 3, scale=(9 * 64) ** -0.5)
    out = ops.conv3x3(x, pack_conv3x3(w).to(DEV))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, 160)
    assert rel(out, ref) < 2e-5
    for C in (64, 192):                            # stride-2 im2col (Downsample3D); 192 = three 64-channel K slabs
        xf = rn(30, 3, 16, 16, C)
        w2 = rn(31, 160, C, 3, 3, scale=(9 * C) ** -0.5)
        cols = ops.im2col3x3(xf.reshape(3, 16, 16, C), stride=2)
        out = ops.gemm(cols, pack_conv3x3(w2).to(DEV))
        ref = F.conv2d(xf.bfloat16().float().permute(0, 3, 1, 2), w2.bfloat16().float(), None, stride=2, padding=1)
        assert rel(out, ref.permute(0, 2, 3, 1).reshape(-1, 160)) < 2e-5


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("B,T,C1,C2", [(2, 3 * 64, 320, 0), (1, 2 * 256, 1280, 640), (2, 48, 640, 320), (3, 1000, 64, 0)])
def test_groupnorm(B, T, C1, C2):
    x1 = rn(40, B * T, C1) * 2 + 0.5
    x2 = rn(41, B * T, C2) - 0.3 if C2 else None
    C = C1 + C2
    g, b = rn(42, C), rn(43, C)
    full = x1 if x2 is None else torch.cat([x1, x2], 1)
    ref = F.group_norm(full.reshape(B, T, C).permute(0, 2, 1).double(), 32, g.double(), b.double(), 1e-5)
    ref = F.silu(ref).permute(0, 2, 1).reshape(B * T, C).float()
    y, raw = ops.groupnorm(x1, x2, B, g, b, 1e-5, True, out_dtype=torch.float32, want_raw=True)
    assert rel(y, ref) < 2e-6
    assert torch.equal(raw, full.bfloat16())
    y16 = ops.groupnorm(x1, x2, B, g, b, 1e-5, True)
    assert y16.dtype == torch.bfloat16 and rel(y16.float(), ref) < 4e-3
    again = ops.groupnorm(x1, x2, B, g, b, 1e-5, True, out_dtype=torch.float32)
    assert torch.equal(again, y)            # deterministic (no float atomics)


@pytest.mark.parametrize("M,C", [(100, 320), (333, 640), (64, 1280)])
def test_layernorm(M, C):
    x = rn(50, M, C) * 3 + 1
    g, b = rn(51, C), rn(52, C)
    ref = F.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5).float()
    y = ops.layernorm(x, g, b)
    assert rel(y.float(), ref) < 4e-3
    assert (y.float() - ref).abs().max() < 0.05 * ref.abs().max()


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, causal):
    s = (q.double() @ k.double().transpose(-1, -2)) * q.shape[-1] ** -0.5
    if causal:
        L = s.shape[-1]
        s = s.masked_fill(~torch.ones(L, L, dtype=torch.bool, device=s.device).tril(), float("-inf"))
    return (s.softmax(-1) @ v.double()).float()


@pytest.mark.parametrize("d,frames,L", [(40, 3, 1024), (80, 2, 256), (160, 5, 64), (160, 3, 16), (40, 1, 100), (40, 5, 256),
                                        (40, 2, 4096), (80, 1, 1024), (80, 3, 100), (80, 40, 256), (80, 2, 64), (160, 40, 64), (160, 2, 256)])
def test_attention_spatial(d, frames, L):
    heads = 8
    C = heads * d
    qkv = rn(60, frames * L, 3 * C).bfloat16()
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SPATIAL, heads=heads, n_outer=frames, Lq=L, Lk=L)
    t = qkv.float().reshape(frames, L, 3, heads, d).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(t[0], t[1], t[2], False).permute(0, 2, 1, 3).reshape(frames * L, C)
    assert rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("mode", ["spatial", "scta"])
def test_attention_tc_lazy_rescale(mode):
    """Scores whose row maximum keeps growing along the key axis (key norms ramp up 6x) and spans tens of log2 units:
    exercises the tcgen05 kernel's lazy-maximum path (O rescaled in TMEM only when a row would exceed 2^8)."""
    heads, d = 8, 40
    C = heads * d
    if mode == "spatial":
        frames, L = 2, 1024
        n = frames * L
        ramp = (1.0 + 5.0 * (torch.arange(n, device=DEV) % L).float() / L)[:, None]
    else:
        B, Fr, H = 1, 8, 32
        n = B * Fr * H * H
        ramp = (1.0 + 5.0 * torch.arange(n, device=DEV).float() / n)[:, None]
    qkv = rn(64, n, 3 * C)
    qkv[:, :C] *= 3.0
    qkv[:, C:2 * C] *= ramp
    qkv = qkv.bfloat16()
    if mode == "spatial":
        out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SPATIAL, heads=heads, n_outer=frames, Lq=L, Lk=L)
        t = qkv.float().reshape(frames, L, 3, heads, d).permute(2, 0, 3, 1, 4)
        ref = _attn_ref(t[0], t[1], t[2], False).permute(0, 2, 1, 3).reshape(n, C)
    else:
        out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=Fr, H=H, W=H)
        t = qkv.float().reshape(B, n // B, 3, heads, d).permute(2, 0, 3, 1, 4)
        ref = torch.empty(B, heads, n // B, d, device=DEV)
        for seq in torch.from_numpy(so.scta_sequences(Fr, H, H)).to(DEV):
            ref[:, :, seq] = _attn_ref(t[0][:, :, seq], t[1][:, :, seq], t[2][:, :, seq], True)
        ref = ref.permute(0, 2, 1, 3).reshape(n, C)
    assert torch.isfinite(out.float()).all()
    assert rel(out.float(), ref) < 8e-3


@pytest.mark.parametrize("d,frames,L", [(40, 3, 1024), (80, 2, 256), (160, 4, 16), (80, 3, 100), (80, 40, 1024), (160, 40, 64)])
def test_attention_cross_77(d, frames, L):
    heads, Lk = 8, 77
    C = heads * d
    q = rn(61, frames * L, C).bfloat16()
    kv = rn(62, frames * Lk, 2 * C).bfloat16()
    out = ops.attention(q, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=frames, Lq=L, Lk=Lk)
    qh = q.float().reshape(frames, L, heads, d).permute(0, 2, 1, 3)
    kh = kv[:, :C].float().reshape(frames, Lk, heads, d).permute(0, 2, 1, 3)
    vh = kv[:, C:].float().reshape(frames, Lk, heads, d).permute(0, 2, 1, 3)
    ref = _attn_ref(qh, kh, vh, False).permute(0, 2, 1, 3).reshape(frames * L, C)
    assert rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("d,B,Fr,H", [(40, 2, 3, 32), (80, 1, 4, 16), (160, 2, 3, 8), (160, 2, 5, 4), (40, 1, 2, 64),
                                      (40, 2, 4, 32), (40, 1, 16, 32), (80, 2, 16, 16), (80, 1, 12, 16), (80, 1, 4, 32),
                                      (80, 3, 8, 16), (80, 1, 3, 16), (160, 2, 16, 8), (160, 1, 12, 8), (160, 2, 16, 4), (160, 1, 7, 4)])
def test_attention_scta(d, B, Fr, H):
    heads = 8
    C = heads * d
    T = Fr * H * H
    qkv = rn(63, B * T, 3 * C).bfloat16()
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=Fr, H=H, W=H)
    t = qkv.float().reshape(B, T, 3, heads, d).permute(2, 0, 3, 1, 4)      # (3, B, heads, T, d)
    ref = torch.empty(B, heads, T, d, device=DEV)
    for seq in torch.from_numpy(so.scta_sequences(Fr, H, H)).to(DEV):
        ref[:, :, seq] = _attn_ref(t[0][:, :, seq], t[1][:, :, seq], t[2][:, :, seq], True)
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, C)
    assert rel(out.float(), ref) < 6e-3


def test_scta_row_permutation_bit_exact(golden_dir):
    """The kernel's gather permutation == the reference's window_partition order (golden from the reference)."""
    g = torch.load(os.path.join(golden_dir, "scta_index.pt"), weights_only=False)
    for key, seqs in g.items():
        if key.startswith("heads"):
            continue
        f, h, w = (int(v) for v in key.split("x"))
        idx = ops.scta_row_index(2, f, h, w).cpu().long()
        assert torch.equal(idx[0], seqs), key
        assert torch.equal(idx[1], seqs + f * h * w), key


# ------------------------------------------------------------------------------------------------ elementwise
def test_rope_matches_oracle():
    heads, d, Fr, hw, B = 8, 40, 3, 64, 2
    C = heads * d
    M = B * Fr * hw
    qkv = rn(70, M, 3 * C).bfloat16()
    freqs = (1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
    ref = qkv.clone().float().reshape(B, Fr * hw, 3, heads, d)
    pos = torch.arange(Fr * hw)
    for which in (0, 1):
        r = so.rope_interleaved(ref[:, :, which].permute(0, 2, 1, 3).cpu(), pos, 32)     # (B, heads, T, d)
        ref[:, :, which] = r.permute(0, 2, 1, 3).to(DEV)
    qkv_tab = qkv.clone()
    ops.rope_inplace(qkv, Fr * hw, heads, d, 0, C, freqs)
    assert rel(qkv.float(), ref.reshape(M, 3 * C)) < 4e-3
    assert torch.equal(qkv[:, 2 * C:].float(), ref.reshape(M, 3 * C)[:, 2 * C:])          # V untouched
    # the vectorised table kernel (fp16 (cos, sin) table): same rotation up to the table's 2^-11 rounding
    ops.rope_inplace(qkv_tab, Fr * hw, heads, d, 0, C, freqs, tab=ops.rope_table(freqs, Fr * hw))
    assert rel(qkv_tab.float(), ref.reshape(M, 3 * C)) < 4e-3
    assert torch.equal(qkv_tab[:, 2 * C:], qkv[:, 2 * C:])


def test_rope_table_kernel_large_grid():
    """More rows than one pass of the capped grid covers: the row loop and the incremental position (pos += step mod T) of the
    table kernel against the direct kernel."""
    heads, d, T, B = 8, 40, 12288, 2
    C = heads * d
    qkv = rn(170, B * T, 3 * C).bfloat16()
    freqs = (1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
    a, b = qkv.clone(), qkv.clone()
    ops.rope_inplace(a, T, heads, d, 0, C, freqs)
    ops.rope_inplace(b, T, heads, d, 0, C, freqs, tab=ops.rope_table(freqs, T))
    assert rel(b.float(), a.float()) < 3e-3
    assert torch.equal(b[:, 2 * C:], qkv[:, 2 * C:])
    # a row far into the second clip: position = row mod T
    row = T + 7777
    x = qkv[row, :32].float().cpu()
    ang = 7777 * freqs.cpu()
    want0 = x[0] * torch.cos(ang[0]) - x[1] * torch.sin(ang[0])
    assert abs(float(b[row, 0]) - float(want0)) < 0.03 * (abs(float(want0)) + 1.0)


def test_time_embedding_and_small_linear():
    t = torch.tensor([991.0, 1.0, 496.0], device=DEV)
    emb = ops.timestep_embedding(t, 320, 0.0, True)
    ref = so.timestep_embedding(t.cpu(), 320, True, 0)
    assert rel(emb, ref) < 1e-5
    w, b, add = rn(71, 1280, 320, scale=320 ** -0.5), rn(72, 1280), rn(73, 1280)
    out = ops.small_linear(emb, w, b, add, silu_in=True, silu_out=True)
    refo = F.silu(F.linear(F.silu(emb), w, b) + add)
    assert rel(out, refo) < 1e-5
    x = rn(74, 19, 1280)
    assert rel(ops.small_linear(x, rn(75, 640, 1280), None), x @ rn(75, 640, 1280).t()) < 1e-5
    # ragged N (not a multiple of the block's 32 columns) and K (not a multiple of the 512-wide shared-memory tile)
    x, w, b = rn(76, 5, 516), rn(77, 1001, 516, scale=516 ** -0.5), rn(78, 1001)
    refr = (x.double() @ w.double().t() + b.double()).float()
    assert rel(ops.small_linear(x, w, b), refr) < 1e-6


def test_conv_in_tensor_core_im2col():
    """conv_in as im2col (36 taps zero-padded to one 64-wide k-block) + tcgen05 GEMM: the bf16-operand product of the same conv."""
    B, Fr, H = 2, 3, 16
    x = rn(180, B, 4, Fr, H, H)
    w, b = rn(181, 320, 4, 3, 3, scale=1 / 6.0), rn(182, 320)
    a = ops.conv_in_im2col(x)
    assert a.shape == (B * Fr * H * H, 64) and a.dtype == torch.bfloat16 and not a[:, 36:].any()
    w64 = torch.zeros(320, 64, device=DEV)
    w64[:, :36] = w.reshape(320, 36)
    r = ops.gemm_ex(a, w64.bfloat16(), bias=b, col_stats=True, out_dtype=torch.bfloat16)
    ref = so.conv_framewise(x.bfloat16().float().cpu(), w.bfloat16().float().cpu(), b.cpu()).permute(0, 2, 3, 4, 1).reshape(-1, 320)
    assert rel(r.out.float(), ref) < 4e-3                       # bf16 rounding of the output only
    exact = so.conv_framewise(x.cpu(), w.cpu(), b.cpu()).permute(0, 2, 3, 4, 1).reshape(-1, 320)
    assert rel(r.out.float(), exact) < 8e-3                     # + bf16 rounding of the latent and the weight
    r32 = ops.gemm_ex(a, w64.bfloat16(), bias=b, col_stats=True)
    assert rel(r32.out, ref) < 2e-5


def test_conv_in_out():
    B, Fr, H = 2, 3, 8
    x = rn(80, B, 4, Fr, H, H)
    w, b = rn(81, 320, 4, 3, 3), rn(82, 320)
    out = ops.conv_in(x, w.reshape(320, 36).contiguous(), b)
    ref = so.conv_framewise(x.cpu(), w.cpu(), b.cpu()).permute(0, 2, 3, 4, 1).reshape(-1, 320)
    assert rel(out, ref) < 1e-5
    # column statistics for the consuming GroupNorms: per 32-row slab (sum, sumsq), same layout as the GEMM's col_stats
    out2, st = ops.conv_in(x, w.reshape(320, 36).contiguous(), b, col_stats=True)
    assert torch.equal(out2, out) and st.shape == (out.shape[0] // 32, 320, 2)
    slabs = out.reshape(-1, 32, 320).double()
    assert rel(st[..., 0], slabs.sum(1)) < 1e-5 and rel(st[..., 1], (slabs * slabs).sum(1)) < 1e-5
    h = rn(83, B * Fr * H * H, 320)
    wo, bo = rn(84, 4, 320, 3, 3, scale=0.02), rn(85, 4)
    from seervideoldm_b200.packing import pack_conv_out
    eps = ops.conv_out(h, pack_conv_out(wo.cpu()).to(DEV), bo, B, Fr, H, H)
    ref = so.conv_framewise(h.cpu().reshape(B, Fr, H, H, 320).permute(0, 4, 1, 2, 3), wo.cpu(), bo.cpu())
    assert eps.shape == (B, 4, Fr, H, H) and rel(eps, ref) < 1e-5
    # row-block kernel geometry: one / two pixel groups per half-warp, ragged last group, fewer output channels, 2 chunks
    for (n_b, n_f, hh, ww, cin, cout) in [(1, 2, 32, 32, 320, 4), (1, 1, 16, 64, 128, 4), (2, 1, 5, 6, 64, 3), (1, 3, 4, 4, 64, 4)]:
        h = rn(86, n_b * n_f * hh * ww, cin)
        wo, bo = rn(87, cout, cin, 3, 3, scale=0.05), rn(88, cout)
        eps = ops.conv_out(h, pack_conv_out(wo.cpu()).to(DEV), bo, n_b, n_f, hh, ww)
        ref = so.conv_framewise(h.cpu().reshape(n_b, n_f, hh, ww, cin).permute(0, 4, 1, 2, 3), wo.cpu(), bo.cpu())
        assert eps.shape == ref.shape and rel(eps, ref) < 1e-5, (n_b, n_f, hh, ww, cin, cout)


def test_upsample_and_cast():
    x = rn(90, 3 * 4 * 4, 64)
    y = ops.upsample2x(x, 3, 4, 4)
    ref = x.reshape(3, 4, 4, 64).repeat_interleave(2, 1).repeat_interleave(2, 2).bfloat16()
    assert torch.equal(y, ref)
    assert torch.equal(ops.cast_bf16(x), x.bfloat16())


def test_cfg_ddim_update_bit_exact():
    """Integer-like exactness: same fp32 operation order as ddim_video.py:209-237 -> bit-identical to PyTorch."""
    b, C, F1, F2, H = 2, 4, 2, 5, 8
    sch = so.make_schedule(30)
    for idx in (30, 17, 0):
        eps = rn(100 + idx, 2 * b, C, F1 + F2, H, H)
        x = rn(200 + idx, b, C, F2, H, H) * 3
        a_t, a_p, s1m = sch.alphas[idx], sch.alphas_prev[idx], sch.sqrt_one_minus_alphas[idx]
        xp, p0 = ops.cfg_ddim_update(eps, x, F1, True, 7.5, float(s1m), float(a_t.sqrt()), float(a_p.sqrt()),
                                     float((1.0 - a_p).sqrt()))
        e_u, e_c = eps.cpu().chunk(2)
        e = so.cfg_combine(e_u[:, :, F1:], e_c[:, :, F1:], 7.5)
        full = lambda v: torch.full((b, 1, 1, 1, 1), float(v))
        xr, pr = so.ddim_update(x.cpu(), e, full(a_t), full(a_p), full(s1m))
        assert torch.equal(p0.cpu(), pr) and torch.equal(xp.cpu(), xr)
