"""State-dict schema of SeerUNet and a deterministic random-init factory.

The schema reproduces the reference's 1006-entry state dict key-for-key and shape-for-shape
(/root/reference/seer/models/unet_3d_condition.py:94-205,
unet_3d_blocks.py:126-208,281-362,433-483,510-588,660-705, attention.py:97-127,181-202,265-279,
429-490,705-742, resnet.py:106-172) so reference checkpoints load with strict=True
(inference.py:126-127).  `random_state_dict` is the synthetic "random-init weights of that
architecture" used by the benchmarks and parity tests: every tensor is drawn from its own
generator seeded by (seed, key), so the same weights can be rebuilt on any machine without
shipping 4.3 GB of fixtures.  `proj_out` weights are drawn N(0, 0.02) instead of the reference's
zero-init so the attention paths contribute to the output.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import UNetConfig

Shape = Tuple[int, ...]


def _resnet(s: "OrderedDict[str, Shape]", p: str, cin: int, cout: int, temb: int) -> None:
    s[p + "norm1.weight"] = (cin,); s[p + "norm1.bias"] = (cin,)
    s[p + "conv1.weight"] = (cout, cin, 3, 3); s[p + "conv1.bias"] = (cout,)
    s[p + "time_emb_proj.weight"] = (cout, temb); s[p + "time_emb_proj.bias"] = (cout,)
    s[p + "norm2.weight"] = (cout,); s[p + "norm2.bias"] = (cout,)
    s[p + "conv2.weight"] = (cout, cout, 3, 3); s[p + "conv2.bias"] = (cout,)
    if cin != cout:
        s[p + "conv_shortcut.weight"] = (cout, cin, 1, 1); s[p + "conv_shortcut.bias"] = (cout,)


def _attn(s, p: str, c: int, ctx: int, bias_out: bool = True, rotary: int = 0) -> None:
    if rotary:
        s[p + "rotary_emb.freqs"] = (rotary // 2,)
    s[p + "to_q.weight"] = (c, c)
    s[p + "to_k.weight"] = (c, ctx)
    s[p + "to_v.weight"] = (c, ctx)
    s[p + "to_out.0.weight"] = (c, c); s[p + "to_out.0.bias"] = (c,)


def _ff(s, p: str, c: int) -> None:
    s[p + "net.0.proj.weight"] = (8 * c, c); s[p + "net.0.proj.bias"] = (8 * c,)
    s[p + "net.2.weight"] = (c, 4 * c); s[p + "net.2.bias"] = (c,)


def _ln(s, p: str, c: int) -> None:
    s[p + "weight"] = (c,); s[p + "bias"] = (c,)


def _transformer(s, p: str, c: int, heads: int, ctx_dim: int, temporal: bool) -> None:
    """SpatialTransformer3D (attention.py:97-127) with its single block."""
    s[p + "norm.weight"] = (c,); s[p + "norm.bias"] = (c,)
    s[p + "proj_in.weight"] = (c, c, 1, 1); s[p + "proj_in.bias"] = (c,)
    b = p + "transformer_blocks.0."
    d = c // heads
    if temporal:     # BasicTransformerBlock3D(temporal=True): attn1 (SCTA), ff, norm1, norm3   (attention.py:184-201)
        _attn(s, b + "attn1.", c, c, rotary=min(32, d))
        _ff(s, b + "ff.", c)
        _ln(s, b + "norm1.", c); _ln(s, b + "norm3.", c)
    else:            # BasicTextTransformerBlock3D: attn1, ff, attn2, norm2, norm1, norm3           (attention.py:268-277)
        _attn(s, b + "attn1.", c, c)
        _ff(s, b + "ff.", c)
        _attn(s, b + "attn2.", c, ctx_dim)
        _ln(s, b + "norm2.", c); _ln(s, b + "norm1.", c); _ln(s, b + "norm3.", c)
    s[p + "proj_out.weight"] = (c, c, 1, 1); s[p + "proj_out.bias"] = (c,)


def unet_schema(cfg: UNetConfig) -> "OrderedDict[str, Shape]":
    """key -> shape, in the reference's registration order."""
    s: "OrderedDict[str, Shape]" = OrderedDict()
    boc = cfg.block_out_channels
    temb, heads, ctx = cfg.time_embed_dim, cfg.heads, cfg.cross_attention_dim
    s["conv_in.weight"] = (boc[0], cfg.in_channels, 3, 3); s["conv_in.bias"] = (boc[0],)
    s["time_embedding.linear_1.weight"] = (temb, boc[0]); s["time_embedding.linear_1.bias"] = (temb,)
    s["time_embedding.linear_2.weight"] = (temb, temb); s["time_embedding.linear_2.bias"] = (temb,)
    n = len(boc)
    out_c = boc[0]
    for i in range(n):
        in_c, out_c = out_c, boc[i]
        p = f"down_blocks.{i}."
        has_attn = i < n - 1
        if has_attn:
            for j in range(cfg.layers_per_block):
                _transformer(s, f"{p}attentions.{j}.", out_c, heads, ctx, False)
            for j in range(cfg.layers_per_block):
                _transformer(s, f"{p}temporal_attentions.{j}.", out_c, heads, ctx, True)
        for j in range(cfg.layers_per_block):
            _resnet(s, f"{p}resnets.{j}.", in_c if j == 0 else out_c, out_c, temb)
        if i < n - 1:
            s[f"{p}downsamplers.0.conv.weight"] = (out_c, out_c, 3, 3); s[f"{p}downsamplers.0.conv.bias"] = (out_c,)
    rev = list(reversed(boc))
    out_c = rev[0]
    for i in range(n):
        prev_c, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, n - 1)]
        p = f"up_blocks.{i}."
        nl = cfg.layers_per_block + 1
        has_attn = i > 0
        if has_attn:
            for j in range(nl):
                _transformer(s, f"{p}attentions.{j}.", out_c, heads, ctx, False)
            for j in range(nl):
                _transformer(s, f"{p}temporal_attentions.{j}.", out_c, heads, ctx, True)
        for j in range(nl):
            skip_c = in_c if j == nl - 1 else out_c
            res_in = prev_c if j == 0 else out_c
            _resnet(s, f"{p}resnets.{j}.", res_in + skip_c, out_c, temb)
        if i < n - 1:
            s[f"{p}upsamplers.0.conv.weight"] = (out_c, out_c, 3, 3); s[f"{p}upsamplers.0.conv.bias"] = (out_c,)
    c = boc[-1]
    _transformer(s, "mid_block.attentions.0.", c, heads, ctx, False)
    _transformer(s, "mid_block.temporal_attentions.0.", c, heads, ctx, True)
    _resnet(s, "mid_block.resnets.0.", c, c, temb)
    _resnet(s, "mid_block.resnets.1.", c, c, temb)
    s["conv_norm_out.weight"] = (boc[0],); s["conv_norm_out.bias"] = (boc[0],)
    s["conv_out.weight"] = (cfg.out_channels, boc[0], 3, 3); s["conv_out.bias"] = (cfg.out_channels,)
    return s


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def random_state_dict(cfg: UNetConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (fp32, CPU).  Linear/conv: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    for weight and bias (PyTorch's default scale); norm weight 1+N(0,0.1), norm bias N(0,0.1);
    proj_out weight N(0, 0.02); rotary freqs are the closed-form buffer."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in unet_schema(cfg).items():
        g = _gen(seed, key)
        if key.endswith("rotary_emb.freqs"):
            dim = 2 * shape[0]
            out[key] = 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
            continue
        owner = key.split(".")[-2]
        is_norm = owner.startswith("norm") or owner == "conv_norm_out"
        if is_norm:
            base = 1.0 if key.endswith("weight") else 0.0
            out[key] = base + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("proj_out.weight"):
            out[key] = 0.02 * torch.randn(shape, generator=g)
        else:
            if key.endswith("weight"):
                fan_in = int(math.prod(shape[1:]))
            else:
                wshape = unet_schema_cache(cfg)[key[: -len("bias")] + "weight"]
                fan_in = int(math.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out


_SCHEMA_CACHE: dict = {}


def unet_schema_cache(cfg: UNetConfig):
    if cfg not in _SCHEMA_CACHE:
        _SCHEMA_CACHE[cfg] = unet_schema(cfg)
    return _SCHEMA_CACHE[cfg]


# ---------------------------------------------------------------------------------------------------------------------
# FSTextTransformer (seer/models/unet_3d_condition.py:379-398; LinearTransformer3D attention.py:152-170, 328-362)
# ---------------------------------------------------------------------------------------------------------------------
MAX_TEXT_LENGTH = 1024    # unet_3d_condition.py MAX_LENGTH: pos_embed is (1, F, 1024, 768), forward slices the first L = 77 tokens


def fstext_schema(num_frames: int = 16, num_layers: int = 2, channels: int = 768, heads: int = 8,
                  cross_attention_dim: int = 768) -> "OrderedDict[str, Shape]":
    """State-dict keys/shapes of the reference FSTextTransformer, in its registration order."""
    s: "OrderedDict[str, Shape]" = OrderedDict()
    c = channels
    s["learnable_query"] = (1, 1, 1, c)
    s["pos_embed"] = (1, num_frames, MAX_TEXT_LENGTH, c)
    for n in range(num_layers):
        b0 = f"trf_blocks.{n}.transformer_blocks.0."
        _attn(s, b0 + "attn1.", c, c)
        _ff(s, b0 + "ff.", c)
        _attn(s, b0 + "attn2.", c, cross_attention_dim)
        _ln(s, b0 + "norm2.", c)
        _ln(s, b0 + "norm1.", c)
        _ln(s, b0 + "norm3.", c)
        b1 = f"trf_blocks.{n}.transformer_blocks.1."
        _attn(s, b1 + "attn1.", c, c, rotary=min(32, c // heads))
        _ff(s, b1 + "ff.", c)
        _ln(s, b1 + "norm1.", c)
        _ln(s, b1 + "norm3.", c)
    _ln(s, "norm.", c)
    return s


def random_fstext_state_dict(num_frames: int = 16, num_layers: int = 2, seed: int = 0, **kw) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic FSText weights (fp32, CPU): linear layers at PyTorch's default scale, norms 1+N(0,0.1) /
    N(0,0.1), learnable_query and pos_embed N(0, 1) / N(0, 0.5) (the reference zero-initialises them; zeros would make
    every token identical and the attention paths degenerate)."""
    schema = fstext_schema(num_frames, num_layers, **kw)
    out: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in schema.items():
        g = _gen(seed, "fstext." + key)
        if key.endswith("rotary_emb.freqs"):
            dim = 2 * shape[0]
            out[key] = 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        elif key == "learnable_query":
            out[key] = torch.randn(shape, generator=g)
        elif key == "pos_embed":
            out[key] = 0.5 * torch.randn(shape, generator=g)
        elif key.split(".")[-2].startswith("norm"):
            out[key] = (1.0 if key.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = shape if key.endswith("weight") else schema[key[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out
