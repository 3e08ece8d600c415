"""Parity at the shapes the BENCHMARK runs (-m gpu, slow): BASELINE.json configs 2-5 at full size.

The bench (bench.py) runs UNet batch 16 x 16 frames x 32x32 latents (M = 262 144 token rows), where the GEMM tile planner
picks plans (`gemm_tc_kernel<160,2>` B-stationary, `<256,2>`, ...) that the small-shape tests never see.  Here:

  * one Bridge-shape CFG evaluation, bf16 path vs the GPU fp32 path (itself pinned <= 1e-4 to the oracle and to golden eps of
    the unmodified reference, tests/test_fp32_path_gpu.py) <= 2e-2, plus the CPU oracle on one clip of the batch (its
    [uncond; cond] pair) — valid because clips never interact;
  * full-size 31-evaluation DDIM+CFG loops (Sthv2 and Bridge shapes), bf16 vs fp32 path <= 5e-2, and one CPU-oracle loop at
    the Sthv2 shape;
  * BASELINE.json config 5 (64x64 latents, 16 frames, batch 4): SCTA / spatial / cross attention kernels at every level
    against an fp64 statement of softmax(QK^T/sqrt d)V;
  * every attention / GEMM launch of a Bridge-shape evaluation must have dispatched to a tcgen05 kernel
    (seer_b200_debug_last_*): a geometry-guard regression would otherwise pass parity on the mma.sync fallback.
Tolerances are north_star's (BASELINE.json): per-step eps rel-L2 <= 2e-2 for bf16, 31-evaluation final latents <= 5e-2."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import seer_oracle as so  # noqa: E402  (checker only)
from seervideoldm_b200 import DDIMSampler, SeerUNet, ops  # noqa: E402
from seervideoldm_b200.config import sd15_config  # noqa: E402
from seervideoldm_b200.pipeline import ddim_sample_latents  # noqa: E402
from seervideoldm_b200.weights import random_state_dict  # noqa: E402

STEP_TOL_BF16 = 2e-2
LOOP_TOL = 5e-2
DEV = "cuda"


def gen(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.fixture(scope="module")
def model():
    cfg = sd15_config(sample_size=32)
    sd = random_state_dict(cfg, seed=0)
    net = SeerUNet(sample_size=32, cross_attention_dim=768)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), sd


def _bridge_inputs(clips=8, frames=16, latent=32):
    x = gen(101, clips, 4, frames, latent, latent)
    c, uc = gen(102, clips, frames, 77, 768), gen(103, clips, 1, 77, 768).expand(-1, frames, -1, -1)
    return torch.cat([x, x]), torch.cat([uc, c]).contiguous()          # the sampler's [uncond; cond] CFG batch


def test_bridge_step_bf16_vs_fp32_path_and_oracle(model):
    """BASELINE.json config 3 (the bench workload): UNet batch 16 (8 clips under CFG), 16 frames, 32x32 latents."""
    net, sd = model
    x, c = _bridge_inputs()
    t = torch.full((16,), 496, dtype=torch.long)
    out16 = net.set_precision("bf16")(x.cuda(), t.cuda(), c.cuda()).cpu()
    out32 = net.set_precision("fp32")(x.cuda(), t.cuda(), c.cuda()).cpu()
    net.set_precision("bf16")
    err = so.rel_l2(out16, out32)
    per_clip = [so.rel_l2(out16[i], out32[i]) for i in range(16)]
    print(f"Bridge-shape step, bf16 vs fp32 path: rel-L2 {err:.3e} (per batch row max {max(per_clip):.3e})")
    assert err < STEP_TOL_BF16 and max(per_clip) < STEP_TOL_BF16
    # CPU oracle on clip 3's [uncond; cond] pair (rows 3 and 11 of the CFG batch): clips never interact
    rows = [3, 11]
    ref = so.unet_forward(sd, x[rows], t[rows], c[rows], 0)
    e32, e16 = so.rel_l2(out32[rows], ref), so.rel_l2(out16[rows], ref)
    print(f"Bridge-shape step, clip 3 vs CPU oracle: fp32 path {e32:.3e}, bf16 path {e16:.3e}")
    assert e32 < 1e-4 and e16 < STEP_TOL_BF16


def _loop(net, precision, b, frames, f1, seed, graph=True):
    net.set_precision(precision)
    f2 = frames - f1
    xT, x0 = gen(seed, b, 4, f2, 32, 32), gen(seed + 1, b, 4, f1, 32, 32)
    c = gen(seed + 2, b, frames, 77, 768)
    uc = gen(seed + 3, b, 1, 77, 768).expand(-1, frames, -1, -1).contiguous()
    sampler = DDIMSampler(torch.device("cuda"), use_cuda_graph=graph)
    lat = ddim_sample_latents(sampler, net, (b, 4, f2, 32, 32), c.cuda(), xT.cuda(), x0.cuda(), ddim_steps=30, scale=7.5, uc=uc.cuda())
    net.set_precision("bf16")
    return lat.cpu(), (xT, x0, c, uc)


@pytest.mark.parametrize("name,b,frames,f1", [("sthv2", 1, 12, 2), ("bridge", 8, 16, 1)])
def test_full_size_ddim_loop_bf16_vs_fp32_path(model, name, b, frames, f1):
    """BASELINE.json configs 2 / 3: 31-evaluation DDIM + CFG 7.5 at full size, bf16 product path vs the fp32 path."""
    net, _ = model
    lat16, _ = _loop(net, "bf16", b, frames, f1, seed=200)
    lat32, _ = _loop(net, "fp32", b, frames, f1, seed=200, graph=False)
    err = so.rel_l2(lat16, lat32)
    per_clip = max(so.rel_l2(lat16[i], lat32[i]) for i in range(b))
    print(f"{name}: 31-evaluation DDIM+CFG final latents, bf16 vs fp32 path: rel-L2 {err:.3e} (worst clip {per_clip:.3e}); "
          f"rms {float(lat32.pow(2).mean().sqrt()):.2f}")
    assert torch.isfinite(lat16).all() and err < LOOP_TOL and per_clip < LOOP_TOL


def test_full_size_sthv2_loop_vs_cpu_oracle(model):
    """The whole Sthv2-shape sampling loop (1 clip, 12 frames / 2 reference frames, 31 CFG evaluations) against the CPU
    oracle — about 3 minutes of host time on the GPU box's 16 cores."""
    net, sd = model
    lat16, (xT, x0, c, uc) = _loop(net, "bf16", 1, 12, 2, seed=300)
    ref, _ = so.ddim_sample_latents(lambda x, t, cc, cf: so.unet_forward(sd, x, t, cc, cf), xT, c, x0, 30, 7.5, uc)
    err = so.rel_l2(lat16, ref)
    print(f"sthv2 full-size loop vs CPU oracle: final-latent rel-L2 {err:.3e}")
    assert err < LOOP_TOL


# ------------------------------------------------------------------------------------------------ config 5: attention stress
def _attn_ref(q, k, v, causal):
    s = (q.double() @ k.double().transpose(-1, -2)) * q.shape[-1] ** -0.5
    if causal:
        L = s.shape[-1]
        s = s.masked_fill(~torch.ones(L, L, dtype=torch.bool, device=s.device).tril(), float("-inf"))
    return (s.softmax(-1) @ v.double()).float()


@pytest.mark.parametrize("h,C", [(64, 320), (32, 640), (16, 1280), (8, 1280)])
def test_config5_scta_and_cross_attention(h, C):
    """BASELINE.json config 5: SCTA over (4, 16, h, h, C) and cross-attention (Lq = h^2, Lk = 77, batch 4*16) at the four
    levels of a 64x64-latent UNet — compared, not just timed."""
    B, Fr, heads = 4, 16, 8
    d = C // heads
    T = Fr * h * h
    g = torch.Generator().manual_seed(400 + h)
    qkv = torch.randn(B * T, 3 * C, generator=g).bfloat16().to(DEV)
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=Fr, H=h, W=h)
    assert "tcgen05" in ops.last_attention_kernel(), ops.last_attention_kernel()
    seqs = torch.from_numpy(so.scta_sequences(Fr, h, h)).to(DEV)
    worst = 0.0
    for b in range(B):                                              # per clip: the fp64 score matrices are 1-2 GB each
        t = qkv[b * T:(b + 1) * T].float().reshape(T, 3, heads, d).permute(1, 2, 0, 3)      # (3, heads, T, d)
        ref = torch.empty(heads, T, d, device=DEV)
        for s0 in range(0, seqs.shape[0], 8):
            sq = seqs[s0:s0 + 8]                                    # (n, L)
            qs, ks, vs = (t[i][:, sq] for i in range(3))            # (heads, n, L, d)
            ref[:, sq.reshape(-1)] = _attn_ref(qs, ks, vs, True).reshape(heads, -1, d)
        ref = ref.permute(1, 0, 2).reshape(T, C)
        worst = max(worst, so.rel_l2(out[b * T:(b + 1) * T].float().cpu(), ref.cpu()))
    print(f"config 5 SCTA h={h} d={d}: L={seqs.shape[1]}, {seqs.shape[0]} windows x {B} clips x {heads} heads, worst-clip rel-L2 {worst:.3e}"
          f" [{ops.last_attention_kernel()}]")
    assert worst < 6e-3
    # FSText cross-attention: Lq = h*h queries per frame against the frame's 77 text tokens
    frames, L, Lk = B * Fr, h * h, 77
    q = torch.randn(frames * L, C, generator=g).bfloat16().to(DEV)
    kv = torch.randn(frames * Lk, 2 * C, generator=g).bfloat16().to(DEV)
    o = ops.attention(q, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=frames, Lq=L, Lk=Lk)
    kern = ops.last_attention_kernel()
    qh = q.float().reshape(frames, L, heads, d).permute(0, 2, 1, 3)
    kh = kv[:, :C].float().reshape(frames, Lk, heads, d).permute(0, 2, 1, 3)
    vh = kv[:, C:].float().reshape(frames, Lk, heads, d).permute(0, 2, 1, 3)
    worst = 0.0
    for f0 in range(0, frames, 8):
        ref = _attn_ref(qh[f0:f0 + 8], kh[f0:f0 + 8], vh[f0:f0 + 8], False).permute(0, 2, 1, 3).reshape(-1, C)
        worst = max(worst, so.rel_l2(o[f0 * L:(f0 + 8) * L].float().cpu(), ref.cpu()))
    print(f"config 5 cross h={h} d={d}: Lq={L} Lk=77 x {frames * heads} problems, rel-L2 {worst:.3e} [{kern}]")
    assert worst < 6e-3 and "tcgen05" in kern


# ------------------------------------------------------------------------------------------------ dispatch assertions
def test_bench_shape_dispatches_to_tcgen05_kernels(model, monkeypatch):
    """Every attention and GEMM / conv launch of one Bridge-shape evaluation runs a tcgen05 kernel: a guard regression in
    seer_b200_attention (silent fallback to the mma.sync kernel) or in the GEMM planner shows up here, not only as lost speed."""
    net, _ = model
    seen_attn, seen_gemm = {}, {}
    real_attention, real_gemm_ex = ops.attention, ops.gemm_ex

    def attention(*a, **k):
        out = real_attention(*a, **k)
        key = (k.get("mode"), a[0].shape[1] // k["heads"], k.get("Lq", 0), k.get("Lk", 0), k.get("H", 0))
        seen_attn[key] = ops.last_attention_kernel()
        return out

    def gemm_ex(*a, **k):
        r = real_gemm_ex(*a, **k)
        if r is not None:
            seen_gemm[(tuple(r.out.shape), a[1].shape[1])] = ops.last_gemm_kernel()
        return r

    monkeypatch.setattr(ops, "attention", attention)
    monkeypatch.setattr(ops, "gemm_ex", gemm_ex)
    x, c = _bridge_inputs()
    net(x.cuda(), torch.full((16,), 250, device=DEV), c.cuda())
    torch.cuda.synchronize()
    assert len(seen_attn) >= 10 and len(seen_gemm) >= 30
    for key, kern in sorted(seen_attn.items(), key=str):
        print("attention", key, "->", kern)
        assert "tcgen05" in kern, (key, kern)
    plans = sorted(set(v.split(" ")[0] for v in seen_gemm.values()))
    print("GEMM plans at the bench shape:", plans)
    for key, kern in seen_gemm.items():
        assert "tcgen05" in kern, (key, kern)
