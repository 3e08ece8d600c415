"""CPU restatement (plain PyTorch, fp32/fp64) of the reference's denoising hot path.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  This is the checker the CUDA path is compared
with on the GPU box, where /root/reference does not exist.  It is written from the operator
specs in SURVEY.md Appendix D, not from the reference's module code, and works on a plain
`state_dict` with the reference's key names.

Pinning: the reference ships no tests or golden vectors, so this file is pinned
against OUTPUTS OF THE REFERENCE ITSELF, run in the build container through oracle/reference_loader.py:
oracle/make_golden.py writes tests/golden/*.pt and tests/test_oracle_golden.py re-checks them
everywhere; tests/test_oracle_vs_reference.py compares directly whenever /root/reference exists.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class OracleCfg:
    """SD-1.5 constants of seer/models/unet_3d_condition.py:64-84 (+ SD's cross_attention_dim=768)."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    heads: int = 8                # `attention_head_dim=8` is used as the HEAD COUNT (unet_3d_blocks.py:170-177)
    groups: int = 32
    norm_eps: float = 1e-5        # resnet + conv_norm_out (unet_3d_condition.py:81,203)
    xf_norm_eps: float = 1e-6     # SpatialTransformer3D.norm (attention.py:109)
    flip_sin_to_cos: bool = True
    freq_shift: int = 0


# ----------------------------------------------------------------------------------------------
# integer / index restatements (must be bit-exact)
# ----------------------------------------------------------------------------------------------
def scta_window_size(h: int) -> int:
    """attention.py:30-33,661-668: 0 means 'one global sequence' (h <= MIN_WIN_SIZE)."""
    if h <= 4:
        return 0
    return 8 if (h // 8) >= 4 else 4


def scta_sequences(f: int, h: int, w: int) -> np.ndarray:
    """Flat token ids (f*h*w + y*w + x) of every SCTA sequence, shape (n_windows, L).

    Window order = (window row, window col); order inside a window = (frame, y in window, x in
    window) — attention.py:42-53 (`permute(2,4,0,1,3,5,6)`), SURVEY F4/Appendix D."""
    ws = scta_window_size(h)
    ids = np.arange(f * h * w, dtype=np.int64).reshape(f, h, w)
    if ws == 0:
        return ids.reshape(1, -1)
    out = []
    for wy in range(h // ws):
        for wx in range(w // ws):
            out.append(ids[:, wy * ws:(wy + 1) * ws, wx * ws:(wx + 1) * ws].reshape(-1))
    return np.stack(out)


def scta_allowed(f: int, h: int, w: int) -> np.ndarray:
    """Boolean (T, T) matrix: query token i may attend key token j  (T = f*h*w).
    Closed form of SURVEY F4: same window and s(j) <= s(i)."""
    T = f * h * w
    allowed = np.zeros((T, T), dtype=bool)
    for seq in scta_sequences(f, h, w):
        L = len(seq)
        tri = np.tril(np.ones((L, L), dtype=bool))
        allowed[np.ix_(seq, seq)] = tri
    return allowed


def ddim_timesteps(num_ddim: int, num_ddpm: int = 1000) -> np.ndarray:
    """ldm/modules/diffusionmodules/util.py:46-60 ('uniform'): range(0, T, T//S) + 1 — 31 entries for S=30."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


# ----------------------------------------------------------------------------------------------
# schedule (ldm/models/diffusion/ddim_video.py:27-68, util.py:21-25,63-74)
# ----------------------------------------------------------------------------------------------
@dataclass
class Schedule:
    timesteps: np.ndarray          # int64 (S',)
    alphas: Tensor                 # fp32 (S',)  a_t
    alphas_prev: Tensor            # fp32 (S',)  a_{t-1}
    sqrt_one_minus_alphas: Tensor  # fp32 (S',)
    sigmas: Tensor                 # fp32 (S',)
    alphas_cumprod: Tensor         # fp32 (1000,)


def make_schedule(num_ddim: int, eta: float = 0.0, num_ddpm: int = 1000,
                  linear_start: float = 1e-4, linear_end: float = 2e-2) -> Schedule:
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_ddpm, dtype=torch.float64) ** 2).numpy()
    acp64 = np.cumprod(1.0 - betas, axis=0)
    acp = torch.tensor(acp64, dtype=torch.float32)
    ts = ddim_timesteps(num_ddim, num_ddpm)
    a = acp[ts]
    a_prev64 = np.asarray([acp[0].item()] + acp[ts[:-1]].tolist())
    sig = eta * np.sqrt((1 - a_prev64) / (1 - a.double().numpy()) * (1 - a.double().numpy() / a_prev64))
    return Schedule(timesteps=ts, alphas=a, alphas_prev=torch.tensor(a_prev64, dtype=torch.float32),
                    sqrt_one_minus_alphas=torch.sqrt(1.0 - a), sigmas=torch.tensor(sig, dtype=torch.float32),
                    alphas_cumprod=acp)


def cfg_combine(e_u: Tensor, e_c: Tensor, scale: float) -> Tensor:
    """ddim_video.py:211."""
    return e_u + scale * (e_c - e_u)


def ddim_update(x: Tensor, e: Tensor, a_t: Tensor, a_prev: Tensor, sqrt_1m_at: Tensor,
                sigma_t: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """ddim_video.py:229-237 in the reference's operation order (fp32)."""
    pred_x0 = (x - sqrt_1m_at * e) / a_t.sqrt()
    if sigma_t is None:
        sigma_t = torch.zeros_like(a_t)
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma_t * noise
    return x_prev, pred_x0


# ----------------------------------------------------------------------------------------------
# operators.  Activations are (B, C, F, H, W) at block boundaries (the
# reference's layout) and (B, F, H, W, C) tokens inside the transformers.
# ----------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, flip_sin_to_cos: bool, shift: float) -> Tensor:
    """diffusers 0.10.2 Timesteps (Appendix B); call site unet_3d_condition.py:307."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - shift))
    arg = t.float()[:, None] * freq[None]
    s, c = torch.sin(arg), torch.cos(arg)
    return torch.cat([c, s], -1) if flip_sin_to_cos else torch.cat([s, c], -1)


def conv_framewise(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1) -> Tensor:
    """InflatedConv3d (unet_3d_condition.py:41-49): Conv2d applied to every frame."""
    B, C, Fr, H, W = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), w, b, stride=stride, padding=w.shape[-1] // 2)
    return y.reshape(B, Fr, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def group_norm_5d(x: Tensor, w: Tensor, b: Tensor, groups: int, eps: float) -> Tensor:
    """GroupNorm over (C/groups, F, H, W) per sample — SURVEY F6 (resnet.py:179,197)."""
    B, C = x.shape[:2]
    xg = x.reshape(B, groups, -1).double()
    mean = xg.mean(-1, keepdim=True)
    var = xg.var(-1, unbiased=False, keepdim=True)
    y = ((xg - mean) / torch.sqrt(var + eps)).to(x.dtype).reshape(x.shape)
    shape = (1, C) + (1,) * (x.dim() - 2)
    return y * w.reshape(shape) + b.reshape(shape)


def resnet_block(sd: Dict[str, Tensor], p: str, x: Tensor, emb: Tensor, cfg: OracleCfg) -> Tensor:
    """ResnetBlock3D.forward, resnet.py:174-208."""
    h = F.silu(group_norm_5d(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.groups, cfg.norm_eps))
    h = conv_framewise(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"])
    t = F.linear(F.silu(emb), sd[p + "time_emb_proj.weight"], sd[p + "time_emb_proj.bias"])
    h = h + t[:, :, None, None, None]
    h = F.silu(group_norm_5d(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.groups, cfg.norm_eps))
    h = conv_framewise(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"])
    if p + "conv_shortcut.weight" in sd:
        x = conv_framewise(x, sd[p + "conv_shortcut.weight"], sd[p + "conv_shortcut.bias"])
    return x + h


def _heads(x: Tensor, n: int) -> Tensor:      # (N, L, C) -> (N, n, L, d)
    N, L, C = x.shape
    return x.reshape(N, L, n, C // n).permute(0, 2, 1, 3)


def _unheads(x: Tensor) -> Tensor:            # (N, n, L, d) -> (N, L, C)
    N, n, L, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(N, L, n * d)


def softmax_attention(q: Tensor, k: Tensor, v: Tensor, causal: bool) -> Tensor:
    """softmax(q k^T / sqrt(d) [+ lower-triangular mask]) v  (attention.py:622-630 + xformers 0.0.13)."""
    d = q.shape[-1]
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    if causal:
        L, Lk = s.shape[-2:]
        keep = torch.ones(L, Lk, dtype=torch.bool).tril()
        s = s.masked_fill(~keep, float("-inf"))
    return torch.matmul(torch.softmax(s, dim=-1), v)


def cross_attention(sd: Dict[str, Tensor], p: str, x: Tensor, ctx: Optional[Tensor], heads: int) -> Tensor:
    """CrossAttention.forward (temporal=False), attention.py:512-554.  x:(N,L,C), ctx:(N,Lk,Cc)|None."""
    src = x if ctx is None else ctx
    q = _heads(F.linear(x, sd[p + "to_q.weight"]), heads)
    k = _heads(F.linear(src, sd[p + "to_k.weight"]), heads)
    v = _heads(F.linear(src, sd[p + "to_v.weight"]), heads)
    o = _unheads(softmax_attention(q, k, v, causal=False))
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def rope_interleaved(x: Tensor, pos: Tensor, rot_dim: int) -> Tensor:
    """rotary-embedding-torch 0.1.5 `rotate_queries_or_keys` (Appendix B): channel pairs (2j,2j+1),
    j < rot_dim/2, rotated by pos * 10000^(-2j/rot_dim); the remaining channels pass through."""
    half = rot_dim // 2
    freqs = 1.0 / (10000.0 ** (torch.arange(0, rot_dim, 2)[:half].float() / rot_dim))
    ang = pos.to(freqs.dtype)[:, None] * freqs[None]            # (L, half)
    cos, sin = ang.cos(), ang.sin()
    xr = x[..., :rot_dim].reshape(*x.shape[:-1], half, 2)
    x0, x1 = xr[..., 0], xr[..., 1]
    y = torch.stack((x0 * cos - x1 * sin, x1 * cos + x0 * sin), dim=-1).reshape(*x.shape[:-1], rot_dim)
    return torch.cat([y, x[..., rot_dim:]], dim=-1)


def scta(sd: Dict[str, Tensor], p: str, x: Tensor, heads: int) -> Tensor:
    """WindowSTempAttention.forward, attention.py:632-703 (xformers path = full lower-triangular
    causality over the (f, wy, wx) window sequence — SURVEY F4).  x:(B,F,H,W,C) -> (B,F,H,W,C)."""
    B, Fr, H, W, C = x.shape
    d = C // heads
    tok = x.reshape(B, Fr * H * W, C)
    q = _heads(F.linear(tok, sd[p + "to_q.weight"]), heads)      # (B, n, T, d)
    k = _heads(F.linear(tok, sd[p + "to_k.weight"]), heads)
    v = _heads(F.linear(tok, sd[p + "to_v.weight"]), heads)
    pos = torch.arange(Fr * H * W)
    rot = min(32, d)
    q, k = rope_interleaved(q, pos, rot), rope_interleaved(k, pos, rot)
    seqs = torch.from_numpy(scta_sequences(Fr, H, W))             # (nWin, L) flat token ids
    o = torch.empty_like(q)
    for seq in seqs:
        o[:, :, seq] = softmax_attention(q[:, :, seq], k[:, :, seq], v[:, :, seq], causal=True)
    out = F.linear(_unheads(o), sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    return out.reshape(B, Fr, H, W, C)


def feed_forward(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """FeedForward/GEGLU, attention.py:744-747,791-793: value half first, gate half second, erf GELU."""
    u = F.linear(x, sd[p + "net.0.proj.weight"], sd[p + "net.0.proj.bias"])
    a, g = u.chunk(2, dim=-1)
    return F.linear(a * F.gelu(g), sd[p + "net.2.weight"], sd[p + "net.2.bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def text_block(sd: Dict[str, Tensor], p: str, x: Tensor, context: Tensor, heads: int) -> Tensor:
    """BasicTextTransformerBlock3D.forward, attention.py:308-327.  x:(B,F,H,W,C); context:(B,F,Lk,Cc)."""
    B, Fr, H, W, C = x.shape
    t = x.reshape(B * Fr, H * W, C)
    t = t + cross_attention(sd, p + "attn1.", _ln(sd, p + "norm1.", t), None, heads)
    ctx = context.reshape(B * Fr, context.shape[-2], context.shape[-1])
    t = t + cross_attention(sd, p + "attn2.", _ln(sd, p + "norm2.", t), ctx, heads)
    t = t + feed_forward(sd, p + "ff.", _ln(sd, p + "norm3.", t))
    return t.reshape(B, Fr, H, W, C)


def temporal_block(sd: Dict[str, Tensor], p: str, x: Tensor, cond_frame: int, heads: int) -> Tensor:
    """BasicTransformerBlock3D.forward (temporal=True), attention.py:231-248."""
    B, Fr, H, W, C = x.shape
    x = x + scta(sd, p + "attn1.", _ln(sd, p + "norm1.", x), heads)
    t = x.reshape(B, Fr, H * W, C)
    upd = t[:, cond_frame:]
    upd = upd + feed_forward(sd, p + "ff.", _ln(sd, p + "norm3.", upd))
    t = torch.cat([t[:, :cond_frame], upd], dim=1)
    return t.reshape(B, Fr, H, W, C)


def spatial_transformer(sd: Dict[str, Tensor], p: str, x: Tensor, context: Optional[Tensor], temporal: bool,
                        cond_frame: int, cfg: OracleCfg) -> Tensor:
    """SpatialTransformer3D.forward, attention.py:129-145.  x:(B,C,F,H,W)."""
    h = group_norm_5d(x, sd[p + "norm.weight"], sd[p + "norm.bias"], cfg.groups, cfg.xf_norm_eps)
    h = conv_framewise(h, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    tok = h.permute(0, 2, 3, 4, 1)
    bp = p + "transformer_blocks.0."
    tok = temporal_block(sd, bp, tok, cond_frame, cfg.heads) if temporal else text_block(sd, bp, tok, context, cfg.heads)
    h = tok.permute(0, 4, 1, 2, 3)
    return conv_framewise(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"]) + x


def upsample_nearest2(x: Tensor) -> Tensor:
    """F.interpolate(scale=(1,2,2), 'nearest'), resnet.py:52."""
    return x.repeat_interleave(2, dim=3).repeat_interleave(2, dim=4)


def unet_forward(sd: Dict[str, Tensor], sample: Tensor, timestep, context: Tensor, cond_frame: int = 0,
                 cfg: OracleCfg = OracleCfg(), taps: Optional[dict] = None) -> Tensor:
    """SeerUNet.forward, unet_3d_condition.py:283-376.  sample:(B,4,F,H,W), context:(B,F,Lk,Cc) -> (B,4,F,H,W).
    `taps`, if given, receives named intermediate activations (for kernel-level parity tests)."""
    sd = {k: v.float() for k, v in sd.items()}
    sample, context = sample.float(), context.float()
    B = sample.shape[0]
    t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep])
    t = t.reshape(-1).broadcast_to((B,)) if t.numel() == 1 else t
    boc = cfg.block_out_channels
    temb = timestep_embedding(t, boc[0], cfg.flip_sin_to_cos, cfg.freq_shift)
    emb = F.linear(F.silu(F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])),
                   sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])

    def tap(name, val):
        if taps is not None:
            taps[name] = val.detach().clone()

    tap("emb", emb)
    x = conv_framewise(sample, sd["conv_in.weight"], sd["conv_in.bias"])
    tap("conv_in", x)
    skips: List[Tensor] = [x]
    nlev = len(boc)
    for i in range(nlev):
        has_attn = i < nlev - 1
        for j in range(cfg.layers_per_block):
            p = f"down_blocks.{i}."
            x = resnet_block(sd, f"{p}resnets.{j}.", x, emb, cfg)
            tap(f"{p}resnets.{j}", x)
            if has_attn:
                x = spatial_transformer(sd, f"{p}attentions.{j}.", x, context, False, cond_frame, cfg)
                tap(f"{p}attentions.{j}", x)
                x = spatial_transformer(sd, f"{p}temporal_attentions.{j}.", x, None, True, cond_frame, cfg)
                tap(f"{p}temporal_attentions.{j}", x)
            skips.append(x)
        if i < nlev - 1:
            x = conv_framewise(x, sd[f"down_blocks.{i}.downsamplers.0.conv.weight"],
                               sd[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
            tap(f"down_blocks.{i}.downsamplers.0", x)
            skips.append(x)
    x = resnet_block(sd, "mid_block.resnets.0.", x, emb, cfg)
    x = spatial_transformer(sd, "mid_block.attentions.0.", x, context, False, cond_frame, cfg)
    x = spatial_transformer(sd, "mid_block.temporal_attentions.0.", x, None, True, cond_frame, cfg)
    x = resnet_block(sd, "mid_block.resnets.1.", x, emb, cfg)
    tap("mid_block", x)
    for i in range(nlev):
        has_attn = i > 0
        p = f"up_blocks.{i}."
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"{p}resnets.{j}.", x, emb, cfg)
            if has_attn:
                x = spatial_transformer(sd, f"{p}attentions.{j}.", x, context, False, cond_frame, cfg)
                x = spatial_transformer(sd, f"{p}temporal_attentions.{j}.", x, None, True, cond_frame, cfg)
            tap(f"{p}{j}", x)
        if i < nlev - 1:
            x = upsample_nearest2(x)
            x = conv_framewise(x, sd[f"{p}upsamplers.0.conv.weight"], sd[f"{p}upsamplers.0.conv.bias"])
            tap(f"{p}upsamplers.0", x)
    x = F.silu(group_norm_5d(x, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], cfg.groups, cfg.norm_eps))
    return conv_framewise(x, sd["conv_out.weight"], sd["conv_out.bias"])


# ----------------------------------------------------------------------------------------------
# sampler loop (ddim_video.py:135-238) — eta = 0 path as the pipelines call it
# ----------------------------------------------------------------------------------------------
def p_sample_ddim(unet_fn, x: Tensor, c: Tensor, t: Tensor, index: int, sch: Schedule, x0_emb: Optional[Tensor],
                  scale: float, uc: Optional[Tensor], cond_frames: int = 0) -> Tuple[Tensor, Tensor]:
    """ddim_video.py:182-238 (is_3d=True)."""
    cond_f = 0
    x_cat = x
    if x0_emb is not None:
        cond_f = x0_emb.shape[2]
        x_cat = torch.cat([x0_emb, x], dim=2)
    if uc is None or scale == 1.0:
        e = unet_fn(x_cat, t, c, 0)[:, :, cond_f:]          # ddim_video.py:196 passes no cond_frame
    else:
        e_u, e_c = unet_fn(torch.cat([x_cat] * 2), torch.cat([t] * 2), torch.cat([uc, c]), cond_frames).chunk(2)
        e = cfg_combine(e_u[:, :, cond_f:], e_c[:, :, cond_f:], scale)
    b = x.shape[0]
    full = lambda v: torch.full((b, 1, 1, 1, 1), float(v), dtype=torch.float32)
    return ddim_update(x, e, full(sch.alphas[index]), full(sch.alphas_prev[index]),
                       full(sch.sqrt_one_minus_alphas[index]), full(sch.sigmas[index]))


def ddim_sample_latents(unet_fn, x_T: Tensor, c: Tensor, x0_emb: Optional[Tensor], ddim_steps: int, scale: float,
                        uc: Optional[Tensor], cond_frames: int = 0):
    """DDIMSampler.sample/ddim_sampling (ddim_video.py:70-180) with eta=0; returns (latents, intermediates)."""
    sch = make_schedule(ddim_steps)
    img = x_T
    total = len(sch.timesteps)
    inter = {"x_inter": [img], "pred_x0": [img]}
    for i, step in enumerate(np.flip(sch.timesteps)):
        index = total - i - 1
        ts = torch.full((x_T.shape[0],), int(step), dtype=torch.long)
        img, pred_x0 = p_sample_ddim(unet_fn, img, c, ts, index, sch, x0_emb, scale, uc, cond_frames)
        if index % 100 == 0 or index == total - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter


def fstext_forward(sd: Dict[str, Tensor], context: Tensor, num_frames: int, heads: int = 8) -> Tensor:
    """FSTextTransformer.forward, seer/models/unet_3d_condition.py:470-484, with LinearTransformer3D /
    BasicLinearTransformerBlock3D (attention.py:172-176, 385-427).  context:(b,L,768) CLIP text embedding ->
    (b,F,L,768) per-frame sub-instruction embedding.  `sd` = the module's state dict.

    Per layer: block 0 (temporal=False) = UNMASKED self-attention over the L tokens of each frame (the causal mask at
    attention.py:521-524 is only built when temporal=True), cross-attention of all F*L queries to the L context tokens
    (3-D context branch, :405-406), GEGLU FF; block 1 (temporal=True) = causal RoPE attention along the frame axis per
    (clip, token) (:393,396,521-530; rot dim min(32, 96), position = frame index), GEGLU FF.  Final LayerNorm (:482)."""
    b, l, c = context.shape
    pos = sd["pos_embed"][:, :, :l, :]
    if sd["pos_embed"].shape[1] != num_frames:      # nearest-neighbour resize along the frame axis (:477-478)
        pos = F.interpolate(pos.permute(0, 3, 1, 2), size=(num_frames, l)).permute(0, 2, 3, 1)
    x = sd["learnable_query"].expand(b, num_frames, l, -1) + pos
    f = num_frames
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("trf_blocks."))
    for n in range(n_layers):
        p0 = f"trf_blocks.{n}.transformer_blocks.0."
        t = x.reshape(b * f, l, c)
        t = t + cross_attention(sd, p0 + "attn1.", _ln(sd, p0 + "norm1.", t), None, heads)
        t = t.reshape(b, f * l, c)
        t = t + cross_attention(sd, p0 + "attn2.", _ln(sd, p0 + "norm2.", t), context, heads)
        t = t + feed_forward(sd, p0 + "ff.", _ln(sd, p0 + "norm3.", t))
        x = t.reshape(b, f, l, c)
        p1 = f"trf_blocks.{n}.transformer_blocks.1."
        t = x.permute(0, 2, 1, 3).reshape(b * l, f, c)
        h = _ln(sd, p1 + "norm1.", t)
        q = _heads(F.linear(h, sd[p1 + "attn1.to_q.weight"]), heads)
        k = _heads(F.linear(h, sd[p1 + "attn1.to_k.weight"]), heads)
        v = _heads(F.linear(h, sd[p1 + "attn1.to_v.weight"]), heads)
        rot = min(32, c // heads)
        q, k = rope_interleaved(q, torch.arange(f), rot), rope_interleaved(k, torch.arange(f), rot)
        o = _unheads(softmax_attention(q, k, v, causal=True))
        t = t + F.linear(o, sd[p1 + "attn1.to_out.0.weight"], sd[p1 + "attn1.to_out.0.bias"])
        t = t + feed_forward(sd, p1 + "ff.", _ln(sd, p1 + "norm3.", t))
        x = t.reshape(b, l, f, c).permute(0, 2, 1, 3)
    return F.layer_norm(x, (c,), sd["norm.weight"], sd["norm.bias"], 1e-5)


def rel_l2(a: Tensor, b: Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
