"""Generate tests/golden/*.pt by running the UNMODIFIED reference (this container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden
The reference has no tests or golden vectors of its own, so these fixtures — outputs
of the reference's own modules on seeded inputs — are what pins oracle/seer_oracle.py.
Weights come from seervideoldm_b200.weights.random_state_dict (rebuilt from the seed, not stored).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader as rl  # noqa: E402
from seervideoldm_b200.config import UNetConfig, sd15_config  # noqa: E402
from seervideoldm_b200.weights import random_state_dict  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NARROW = dict(block_out_channels=(64, 128, 128, 128), cross_attention_dim=64)


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def fstext_golden(ref):
    """FSTextTransformer (seer/models/unet_3d_condition.py:379-484) on seeded weights / contexts.  The full outputs are
    (b, F, 77, 768); the fixture keeps every 7th token and every 8th channel of each (plus mean / std of the whole)."""
    from seervideoldm_b200.weights import random_fstext_state_dict
    sd = random_fstext_state_dict(num_frames=16, num_layers=2, seed=0)
    m = ref.FSTextTransformer(num_frames=16, num_layers=2).eval()
    rl.enable_xformers_path(m)
    m.load_state_dict(sd, strict=True)
    cases = []
    for i, (b, nf) in enumerate([(2, 16), (1, 6), (3, 12)]):
        ctx = randn(7000 + i, b, 77, 768)
        m.set_numframe(nf)
        y = m(context=ctx)
        cases.append(dict(ctx_seed=7000 + i, b=b, num_frames=nf, y_sub=y[:, :, ::7, ::8].clone(), mean=float(y.mean()),
                          std=float(y.std())))
    torch.save(dict(weight_seed=0, num_frames=16, num_layers=2, token_stride=7, channel_stride=8, cases=cases),
               os.path.join(OUT, "fstext.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = rl.load()
    torch.set_grad_enabled(False)
    if len(sys.argv) > 1 and sys.argv[1] == "fstext":       # regenerate only the FSText fixture
        fstext_golden(ref)
        return
    fstext_golden(ref)

    # ---- schedule (ldm/models/diffusion/ddim_video.py:27-68) -------------------------------
    for S in (30, 10):
        s = ref.DDIMSampler("cpu")
        s.make_schedule(S, ddim_eta=0.0, verbose=False)
        torch.save(dict(S=S, timesteps=torch.from_numpy(np.asarray(s.ddim_timesteps)).long(),
                        alphas=s.ddim_alphas.float(), alphas_prev=torch.tensor(np.asarray(s.ddim_alphas_prev)),
                        sqrt_one_minus_alphas=torch.as_tensor(np.asarray(s.ddim_sqrt_one_minus_alphas)).float(),
                        sigmas=torch.as_tensor(np.asarray(s.ddim_sigmas)).float(),
                        alphas_cumprod_fp32=s.alphas_cumprod.float()),
                   os.path.join(OUT, f"schedule_{S}.pt"))

    # ---- SCTA index order (seer/models/attention.py:42-69, 661-680) -------------------------
    idx = {}
    for (f, h, w) in [(3, 8, 8), (2, 4, 4), (2, 16, 16), (2, 32, 32), (2, 64, 64), (5, 2, 2)]:
        ids = torch.arange(f * h * w).reshape(1, f, h, w, 1)
        if h > ref.attention.MIN_WIN_SIZE:
            ws = ref.attention.MAX_WIN_SIZE if (h // ref.attention.MAX_WIN_SIZE) >= ref.attention.MAX_RATIO \
                else ref.attention.MIN_WIN_SIZE
            seqs = ref.attention.window_partition(ids, ws)[..., 0]
            back = ref.attention.window_reverse(seqs[..., None], ws, f, h, w)[0, :, 0]
            assert torch.equal(back, torch.arange(f * h * w))
        else:
            seqs = ids.reshape(1, -1)
        idx[f"{f}x{h}x{w}"] = seqs.long()
    # head split/merge order (attention.py:492-504): batch index = b*heads + head
    cross = ref.attention.CrossAttention(query_dim=16, heads=4, dim_head=4)
    t = torch.arange(2 * 3 * 16).reshape(2, 3, 16).float()
    idx["heads_to_batch_2x3x16_h4"] = cross.reshape_heads_to_batch_dim(t).long()
    torch.save(idx, os.path.join(OUT, "scta_index.pt"))

    # ---- SCTA causal dependency pattern, measured on the reference by autograd ---------------
    dep = {}
    for (f, h, w) in [(3, 8, 8), (2, 4, 4)]:
        torch.manual_seed(0)
        attn = ref.attention.WindowSTempAttention(query_dim=32, heads=2, dim_head=16, temporal=True, causal=True)
        attn._use_memory_efficient_attention_xformers = True
        with torch.enable_grad():
            x = torch.randn(1, f, h, w, 32, requires_grad=True)
            y = attn(x)                                   # (1, f*h*w, 32)
            T = f * h * w
            m = torch.zeros(T, T, dtype=torch.bool)
            for i in range(T):
                g, = torch.autograd.grad(y[0, i].sum(), x, retain_graph=True)
                m[i] = g.reshape(T, -1).abs().sum(-1) > 0
        dep[f"{f}x{h}x{w}"] = m
    torch.save(dep, os.path.join(OUT, "scta_dependency.pt"))

    # ---- module level: WindowSTempAttention / transformer blocks on narrow widths ------------
    mods = {}
    for h in (4, 8, 16, 32):
        torch.manual_seed(h)
        C, f = 64, 3
        attn = ref.attention.WindowSTempAttention(query_dim=C, heads=8, dim_head=C // 8, temporal=True, causal=True)
        attn._use_memory_efficient_attention_xformers = True
        x = randn(100 + h, 2, f, h, h, C)
        mods[f"scta_h{h}"] = dict(sd={k: v.clone() for k, v in attn.state_dict().items()}, x=x,
                                  y=attn(x).reshape(2, f, h, h, C))
    for cf in (0, 1, 2):
        torch.manual_seed(50 + cf)
        st = ref.attention.SpatialTransformer3D(64, 8, 8, depth=1, context_dim=None, temporal=True, causal=True)
        rl.enable_xformers_path(st)
        st.proj_out.weight.normal_(std=0.05)
        x = randn(200 + cf, 1, 64, 4, 8, 8)
        mods[f"temporal_xf_cf{cf}"] = dict(sd={k: v.clone() for k, v in st.state_dict().items()}, x=x,
                                           y=st(x, cond_frame=cf), cond_frame=cf)
    torch.manual_seed(60)
    st = ref.attention.SpatialTransformer3D(64, 8, 8, depth=1, context_dim=48, text_frame_condition=True)
    rl.enable_xformers_path(st)
    st.proj_out.weight.normal_(std=0.05)
    x, c = randn(300, 2, 64, 3, 8, 8), randn(301, 2, 3, 77, 48)
    mods["text_xf"] = dict(sd={k: v.clone() for k, v in st.state_dict().items()}, x=x, c=c, y=st(x, context=c))
    torch.manual_seed(61)
    rb = ref.resnet.ResnetBlock3D(in_channels=96, out_channels=64, temb_channels=128, eps=1e-5, non_linearity="silu")
    x, e = randn(302, 2, 96, 3, 8, 8), randn(303, 2, 128)
    mods["resnet_96_64"] = dict(sd={k: v.clone() for k, v in rb.state_dict().items()}, x=x, emb=e, y=rb(x, e))
    torch.save(mods, os.path.join(OUT, "modules.pt"))

    # ---- UNet forward, narrow config ----------------------------------------------------------
    cfg = UNetConfig(sample_size=32, **NARROW)
    sd = random_state_dict(cfg, seed=0)
    net = rl.build_unet(ref, **NARROW)
    net.load_state_dict(sd, strict=True)
    cases = []
    for i, (B, Fr, H, cf, tval) in enumerate([(1, 3, 16, 0, 991), (2, 4, 32, 0, 496), (1, 2, 8, 1, 1), (1, 12, 32, 0, 991)]):
        x, c = randn(1000 + i, B, 4, Fr, H, H), randn(2000 + i, B, Fr, 77, 64)
        t = torch.full((B,), tval, dtype=torch.long)
        cases.append(dict(x=x, c=c, t=t, cond_frame=cf, y=net(x, t, c, cond_frame=cf)))
    torch.save(dict(cfg=NARROW, weight_seed=0, cases=cases), os.path.join(OUT, "unet_narrow.pt"))

    # ---- 31-evaluation DDIM+CFG loop, narrow config (ddim_video.py:70-238) ----------------------
    s = ref.DDIMSampler("cpu")
    b, F1, F2, H = 1, 2, 4, 16
    xT, x0 = randn(3000, b, 4, F2, H, H), randn(3001, b, 4, F1, H, H)
    c = randn(3002, b, F1 + F2, 77, 64)
    uc = randn(3003, b, 1, 77, 64).expand(-1, F1 + F2, -1, -1).contiguous()
    calls = []
    def spy(x_in, t_in, c_in, cond_frame=0):
        calls.append((tuple(x_in.shape), t_in.tolist(), cond_frame))
        return net(x_in, t_in, c_in, cond_frame=cond_frame)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        lat, inter = s.sample(unet=spy, S=30, conditioning=c, batch_size=b, shape=(4, F2, H, H), x0_emb=x0,
                              verbose=False, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                              eta=0.0, x_T=xT, is_3d=True)
    torch.save(dict(cfg=NARROW, weight_seed=0, xT=xT, x0=x0, c=c, uc=uc, scale=7.5, S=30, latents=lat,
                    x_inter=[t.clone() for t in inter["x_inter"]], pred_x0=[t.clone() for t in inter["pred_x0"]],
                    calls=calls), os.path.join(OUT, "ddim_loop_narrow.pt"))

    # ---- UNet forward, full SD-1.5 width, tiny clip (keeps the fixture small) ---------------------
    cfg = sd15_config(sample_size=32)
    sd = random_state_dict(cfg, seed=0)
    net = rl.build_unet(ref)
    net.load_state_dict(sd, strict=True)
    cases = []
    for i, (B, Fr, H, cf, tval) in enumerate([(1, 2, 8, 0, 991), (2, 3, 16, 0, 496)]):
        x, c = randn(4000 + i, B, 4, Fr, H, H), randn(5000 + i, B, Fr, 77, 768)
        t = torch.full((B,), tval, dtype=torch.long)
        cases.append(dict(x_seed=4000 + i, c_seed=5000 + i, shape=(B, Fr, H), t=t, cond_frame=cf,
                          y=net(x, t, c, cond_frame=cf)))
    torch.save(dict(weight_seed=0, cases=cases), os.path.join(OUT, "unet_sd15.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
