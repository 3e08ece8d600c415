"""fp32-parity forward of SeerUNet (north-star tolerance: per-step noise prediction within rel-L2 1e-4 of the
reference's fp32 PyTorch forward, /root/reference/seer/models/unet_3d_condition.py:283-376).

Same kernels as the bf16 product path where those are already fp32 (conv_in / conv_out, time embedding, GroupNorm,
CFG + DDIM update); every contraction runs on the SAME tcgen05 GEMM / implicit-GEMM conv kernel with
error-compensated bf16 operand pairs (csrc/fp32_path.cu, packing.split3_weight):

    A' = [a_hi | a_hi | a_lo],  W' = [w_hi | w_lo | w_hi]   ->   A'.W' = a_hi.w_hi + a_hi.w_lo + a_lo.w_hi   (fp32 accumulate)

and everything between the GEMMs (LayerNorm, GEGLU with exact erf, RoPE, softmax attention) stays fp32 in HBM.  Nothing is
fused here on purpose: this is the parity mode, `SeerUNet.set_precision("bf16")` (default) is the product.
No CPU / PyTorch fallback: every op is a seer_b200 kernel.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops, packing


def _w3_linear(w: torch.Tensor) -> torch.Tensor:
    return packing.pack_linear(packing.split3_weight(w))


def _w3_conv1x1(w: torch.Tensor) -> torch.Tensor:
    return packing.pack_linear(packing.split3_weight(w.reshape(w.shape[0], w.shape[1])))


def _w3_conv3x3(w: torch.Tensor, shortcut: Optional[torch.Tensor] = None) -> torch.Tensor:
    sc = None if shortcut is None else packing.split3_weight(shortcut.reshape(shortcut.shape[0], shortcut.shape[1]))
    return packing.pack_conv3x3(packing.split3_weight(w), sc)


@torch.no_grad()
def pack_fp32(net) -> dict:
    """Split-operand weight pack of the whole UNet (built once, on first fp32 forward)."""
    P = net._P
    f32 = lambda k: P(k).float().contiguous()
    cfg = net.cfg
    pk: dict = {}
    pk["conv_in_w"] = P("conv_in.weight").float().reshape(cfg.block_out_channels[0], -1).contiguous()
    pk["conv_in_b"] = f32("conv_in.bias")
    pk["te1_w"], pk["te1_b"] = f32("time_embedding.linear_1.weight"), f32("time_embedding.linear_1.bias")
    pk["te2_w"], pk["te2_b"] = f32("time_embedding.linear_2.weight"), f32("time_embedding.linear_2.bias")
    pk["gno_g"], pk["gno_b"] = f32("conv_norm_out.weight"), f32("conv_norm_out.bias")
    pk["conv_out_w"] = packing.pack_conv_out(P("conv_out.weight"))
    pk["conv_out_b"] = f32("conv_out.bias")
    temb_w, temb_b = [], []
    off = 0

    def resnet(prefix: str) -> dict:
        nonlocal off
        w1 = P(prefix + "conv1.weight")
        cout, cin = w1.shape[:2]
        r = dict(cin=cin, cout=cout, off=off, sc=net._has(prefix + "conv_shortcut.weight"))
        r["g1"], r["b1"] = f32(prefix + "norm1.weight"), f32(prefix + "norm1.bias")
        r["g2"], r["b2"] = f32(prefix + "norm2.weight"), f32(prefix + "norm2.bias")
        r["w1"] = _w3_conv3x3(w1)
        temb_w.append(f32(prefix + "time_emb_proj.weight"))
        temb_b.append(f32(prefix + "time_emb_proj.bias") + f32(prefix + "conv1.bias"))
        off += cout
        if r["sc"]:
            r["w2"] = _w3_conv3x3(P(prefix + "conv2.weight"), P(prefix + "conv_shortcut.weight"))
            r["bias2"] = f32(prefix + "conv2.bias") + f32(prefix + "conv_shortcut.bias")
        else:
            r["w2"] = _w3_conv3x3(P(prefix + "conv2.weight"))
            r["bias2"] = f32(prefix + "conv2.bias")
        return r

    def xf(prefix: str, temporal: bool) -> dict:
        b = prefix + "transformer_blocks.0."
        t = dict(temporal=temporal, C=P(prefix + "proj_in.weight").shape[0])
        t["gn_g"], t["gn_b"] = f32(prefix + "norm.weight"), f32(prefix + "norm.bias")
        t["pin_w"], t["pin_b"] = _w3_conv1x1(P(prefix + "proj_in.weight")), f32(prefix + "proj_in.bias")
        t["pout_w"], t["pout_b"] = _w3_conv1x1(P(prefix + "proj_out.weight")), f32(prefix + "proj_out.bias")
        for i in (1, 3) if temporal else (1, 2, 3):
            t[f"ln{i}_g"], t[f"ln{i}_b"] = f32(b + f"norm{i}.weight"), f32(b + f"norm{i}.bias")
        t["qkv_w"] = _w3_linear(torch.cat([P(b + "attn1.to_q.weight"), P(b + "attn1.to_k.weight"), P(b + "attn1.to_v.weight")], 0))
        t["o1_w"], t["o1_b"] = _w3_linear(P(b + "attn1.to_out.0.weight")), f32(b + "attn1.to_out.0.bias")
        t["ff1_w"], t["ff1_b"] = _w3_linear(P(b + "ff.net.0.proj.weight")), f32(b + "ff.net.0.proj.bias")
        t["ff2_w"], t["ff2_b"] = _w3_linear(P(b + "ff.net.2.weight")), f32(b + "ff.net.2.bias")
        if temporal:
            t["freqs"] = f32(b + "attn1.rotary_emb.freqs")
        else:
            t["q2_w"] = _w3_linear(P(b + "attn2.to_q.weight"))
            t["kv2_w"] = _w3_linear(torch.cat([P(b + "attn2.to_k.weight"), P(b + "attn2.to_v.weight")], 0))
            t["o2_w"], t["o2_b"] = _w3_linear(P(b + "attn2.to_out.0.weight")), f32(b + "attn2.to_out.0.bias")
        return t

    n = len(cfg.block_out_channels)
    L = cfg.layers_per_block
    pk["down"] = []
    for i in range(n):
        blk = dict(res=[], attn=[], tattn=[], down=None)
        for j in range(L):
            blk["res"].append(resnet(f"down_blocks.{i}.resnets.{j}."))
            if i < n - 1:
                blk["attn"].append(xf(f"down_blocks.{i}.attentions.{j}.", False))
                blk["tattn"].append(xf(f"down_blocks.{i}.temporal_attentions.{j}.", True))
        if i < n - 1:
            blk["down"] = (_w3_conv3x3(P(f"down_blocks.{i}.downsamplers.0.conv.weight")), f32(f"down_blocks.{i}.downsamplers.0.conv.bias"))
        pk["down"].append(blk)
    pk["mid"] = dict(res=[resnet("mid_block.resnets.0."), resnet("mid_block.resnets.1.")],
                     attn=xf("mid_block.attentions.0.", False), tattn=xf("mid_block.temporal_attentions.0.", True))
    pk["up"] = []
    for i in range(n):
        blk = dict(res=[], attn=[], tattn=[], up=None)
        for j in range(L + 1):
            blk["res"].append(resnet(f"up_blocks.{i}.resnets.{j}."))
            if i > 0:
                blk["attn"].append(xf(f"up_blocks.{i}.attentions.{j}.", False))
                blk["tattn"].append(xf(f"up_blocks.{i}.temporal_attentions.{j}.", True))
        if i < n - 1:
            blk["up"] = (_w3_conv3x3(P(f"up_blocks.{i}.upsamplers.0.conv.weight")), f32(f"up_blocks.{i}.upsamplers.0.conv.bias"))
        pk["up"].append(blk)
    pk["temb_w"] = torch.cat(temb_w, 0).contiguous()
    pk["temb_b"] = torch.cat(temb_b, 0).contiguous()
    return pk


# ---------------------------------------------------------------------------------------------------------------------
def linear32(x: torch.Tensor, w3: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [M, K] @ W^T (+bias, +fp32 residual) -> fp32, through the split-operand tcgen05 GEMM."""
    return ops.gemm_ex(ops.split3(x), w3, bias=bias, residual=residual, out=out).out


def conv3x3_32(x: torch.Tensor, n_img: int, H: int, W: int, w3: torch.Tensor, **kw) -> torch.Tensor:
    """x fp32 [n_img*H*W, C] -> fp32 [n_img*H*W, Cout] (frame-wise 3x3, pad 1)."""
    C = x.shape[1]
    return ops.conv3x3_ex(ops.split3(x).view(n_img, H, W, 3 * C), w3, **kw).out


def _resnet(net, r: dict, x1, x2, B, F, H, W, temb_all):
    """ResnetBlock3D.forward (resnet.py:174-208) on the virtual concat [x1 | x2], fp32."""
    T = F * H * W
    eps = net.cfg.norm_eps
    cin, cout = r["cin"], r["cout"]
    h = ops.groupnorm(x1, x2, B, r["g1"], r["b1"], eps, True, out_dtype=torch.float32)
    tb = temb_all[:, r["off"]: r["off"] + cout]
    c1 = conv3x3_32(h, B * F, H, W, r["w1"], bias=tb, bias_div=T)
    h2 = ops.groupnorm(c1, None, B, r["g2"], r["b2"], eps, True, out_dtype=torch.float32)
    if r["sc"]:
        raw = torch.empty((x1.shape[0], 3 * cin), device=x1.device, dtype=torch.bfloat16)
        ops.split3(x1, out=raw, col0=0, ctot=cin)
        if x2 is not None:
            ops.split3(x2, out=raw, col0=x1.shape[1], ctot=cin)
        return conv3x3_32(h2, B * F, H, W, r["w2"], a2=raw, bias=r["bias2"])
    if x2 is not None:
        raise RuntimeError("concat input without a shortcut conv cannot occur in this architecture")
    return conv3x3_32(h2, B * F, H, W, r["w2"], bias=r["bias2"], residual=x1)


def _ff(t: dict, tok: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x + FF(LN3(x)) (attention.py:244,323 + 744-747,791-793), fp32."""
    n3 = ops.layernorm_f32(tok, t["ln3_g"], t["ln3_b"])
    g = ops.geglu_f32(linear32(n3, t["ff1_w"], t["ff1_b"]))
    return linear32(g, t["ff2_w"], t["ff2_b"], residual=tok, out=out)


def _transformer(net, t: dict, xt, B, F, H, W, kv, cond_frame):
    """SpatialTransformer3D.forward (attention.py:129-145) with its text (:308-327) or temporal (:231-248) block, fp32."""
    C, heads = t["C"], net.cfg.heads
    d = C // heads
    hw, T = H * W, F * H * W
    M = B * T
    hn = ops.groupnorm(xt, None, B, t["gn_g"], t["gn_b"], 1e-6, False, out_dtype=torch.float32)
    tok = linear32(hn, t["pin_w"], t["pin_b"])
    qkv = linear32(ops.layernorm_f32(tok, t["ln1_g"], t["ln1_b"]), t["qkv_w"])
    if t["temporal"]:
        ops.rope_ex(qkv, 1, T, heads, d, 0, C, t["freqs"])
        att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=F, H=H, W=W)
    else:
        att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SPATIAL, heads=heads, n_outer=B * F,
                            Lq=hw, Lk=hw)
    tok = linear32(att, t["o1_w"], t["o1_b"], residual=tok)
    if not t["temporal"]:
        q2 = linear32(ops.layernorm_f32(tok, t["ln2_g"], t["ln2_b"]), t["q2_w"])
        Lk = kv.shape[0] // (B * F)
        att2 = ops.attention(q2, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=B * F, Lq=hw, Lk=Lk)
        tok = linear32(att2, t["o2_w"], t["o2_b"], residual=tok)
    if t["temporal"] and cond_frame > 0:
        # the first cond_frame frames of every clip bypass the feed-forward (attention.py:240-246)
        y = tok.clone()
        c0 = min(cond_frame, F) * hw
        for b in range(B):
            mid, hi = b * T + c0, (b + 1) * T
            if mid < hi:
                _ff(t, tok[mid:hi], out=y[mid:hi])
    else:
        y = _ff(t, tok)
    return linear32(y, t["pout_w"], t["pout_b"], residual=xt)


def context_kv(net, pk: dict, context: torch.Tensor, out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """fp32 K/V projections of the text context for every cross-attention layer ([B*F*L, 2C] each, attention.py:517-518)."""
    ctx3 = ops.split3(context.reshape(-1, context.shape[-1]).float().contiguous())
    layers = [a for blk in pk["down"] for a in blk["attn"]] + [pk["mid"]["attn"]] + [a for blk in pk["up"] for a in blk["attn"]]
    if out is None:
        return [ops.gemm_ex(ctx3, a["kv2_w"]).out for a in layers]
    for a, o in zip(layers, out):
        ops.gemm_ex(ctx3, a["kv2_w"], out=o)
    return out


@torch.no_grad()
def forward(net, pk: dict, sample: torch.Tensor, t: torch.Tensor, context: torch.Tensor, cond_frame: int,
            kv: List[torch.Tensor]) -> torch.Tensor:
    """Same sequencing as SeerUNet.forward's bf16 path (unet.py), one fp32 tensor per activation."""
    cfg = net.cfg
    B, _, F, H, W = sample.shape
    temb = ops.timestep_embedding(t, cfg.block_out_channels[0], float(cfg.freq_shift), cfg.flip_sin_to_cos)
    e1 = ops.small_linear(temb, pk["te1_w"], pk["te1_b"], silu_out=True)
    emb = ops.small_linear(e1, pk["te2_w"], pk["te2_b"], silu_out=True)   # = SiLU(emb): emb is consumed only through time_emb_proj(SiLU(emb)), resnet.py:190-192
    temb_all = ops.small_linear(emb, pk["temb_w"], pk["temb_b"])
    kvs = iter(kv)
    x = ops.conv_in(sample.contiguous(), pk["conv_in_w"], pk["conv_in_b"])
    h, w = H, W
    skips = [x]
    for blk in pk["down"]:
        for j, r in enumerate(blk["res"]):
            x = _resnet(net, r, x, None, B, F, h, w, temb_all)
            if blk["attn"]:
                x = _transformer(net, blk["attn"][j], x, B, F, h, w, next(kvs), cond_frame)
                x = _transformer(net, blk["tattn"][j], x, B, F, h, w, None, cond_frame)
            skips.append(x)
        if blk["down"] is not None:
            wd, bd = blk["down"]
            C = x.shape[1]
            cols = ops.im2col3x3(ops.split3(x).view(B * F, h, w, 3 * C), stride=2)
            x = ops.gemm_ex(cols, wd, bias=bd).out
            h, w = h // 2, w // 2
            skips.append(x)
    m = pk["mid"]
    x = _resnet(net, m["res"][0], x, None, B, F, h, w, temb_all)
    x = _transformer(net, m["attn"], x, B, F, h, w, next(kvs), cond_frame)
    x = _transformer(net, m["tattn"], x, B, F, h, w, None, cond_frame)
    x = _resnet(net, m["res"][1], x, None, B, F, h, w, temb_all)
    for blk in pk["up"]:
        for j, r in enumerate(blk["res"]):
            x = _resnet(net, r, x, skips.pop(), B, F, h, w, temb_all)
            if blk["attn"]:
                x = _transformer(net, blk["attn"][j], x, B, F, h, w, next(kvs), cond_frame)
                x = _transformer(net, blk["tattn"][j], x, B, F, h, w, None, cond_frame)
        if blk["up"] is not None:
            wu, bu = blk["up"]
            C = x.shape[1]
            up = ops.split3(x, upsample=(B * F, h, w)).view(B * F, 2 * h, 2 * w, 3 * C)
            h, w = 2 * h, 2 * w
            x = ops.conv3x3_ex(up, wu, bias=bu).out
    y = ops.groupnorm(x, None, B, pk["gno_g"], pk["gno_b"], cfg.norm_eps, True, out_dtype=torch.float32)
    return ops.conv_out(y, pk["conv_out_w"], pk["conv_out_b"], B, F, h, w)
