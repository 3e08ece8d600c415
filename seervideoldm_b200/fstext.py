"""FSTextTransformer — the FSText sub-instruction decomposer that feeds the UNet's cross-attention (SURVEY §8 a18 / §8f
rank 1), over the seer_b200 kernels.

Drop-in for /root/reference/seer/models/unet_3d_condition.py:379-484 (with LinearTransformer3D /
BasicLinearTransformerBlock3D, attention.py:152-176, 328-427): same constructor arguments, `set_numframe`,
`forward(context) -> (b, F, L, 768)`, same state-dict keys / shapes (loads `pytorch_model_1.bin` with strict=True,
inference.py:123-124).

Token stream: fp32 [(b f l), 768] in the reference's (b, f, l) order throughout.  The reference's three reshapes are
index functions of the attention kernel, no copies:
  * block 0 self-attention over the L tokens of a frame   -> SEER_ATTN_SPATIAL, n_outer = b*F, Lq = Lk = L
  * block 0 cross-attention of F*L queries to L CLIP tokens -> SEER_ATTN_CROSS,  n_outer = b,   Lq = F*L, Lk = L
  * block 1 causal RoPE attention along the frame axis     -> SEER_ATTN_FRAME,  n_outer = b, F frames, H = L tokens
    (the reference permutes to (b l) f c and back, attention.py:393,415)
Precision follows `set_precision`: "bf16" = bf16 tensor-core operands with fp32 accumulation / residual stream,
"fp32" = the error-compensated split-operand GEMMs and fp32 kernels of unet_fp32.py.  No CPU / PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops, packing
from .unet import _Node
from .weights import fstext_schema


class FSTextTransformer(nn.Module):
    def __init__(self, num_frames=None, in_channels=768, out_channels=768, n_heads=8, num_layers=2, cross_attention_dim=768):
        super().__init__()
        if in_channels != out_channels:
            raise ValueError("FSTextTransformer: in_channels != out_channels is never used by the reference pipelines")
        if num_frames is None:
            raise ValueError("num_frames is required (size of the learned frame position table)")
        self.num_frames = num_frames
        self.channels, self.heads, self.num_layers = out_channels, n_heads, num_layers
        self.cross_attention_dim = cross_attention_dim
        self.precision = "bf16"
        self._packed = None
        for key, shape in fstext_schema(num_frames, num_layers, out_channels, n_heads, cross_attention_dim).items():
            parts = key.split(".")
            mod: nn.Module = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Node())
                mod = mod._modules[p]
            if key.endswith("rotary_emb.freqs"):
                dim = 2 * shape[0]
                mod.register_buffer("freqs", 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)))
            else:
                mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self) -> None:
        """PyTorch-default-scale init of the linear layers; query / position table zero like the reference (:392-393)."""
        params = dict(self.named_parameters())
        for name, p in params.items():
            if name in ("learnable_query", "pos_embed"):
                p.zero_()
            elif name.split(".")[-2].startswith("norm"):
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            else:
                ref = p if name.endswith("weight") else params[name[: -len("bias")] + "weight"]
                bound = 1.0 / max(1, ref[0].numel()) ** 0.5
                p.uniform_(-bound, bound)
        self._packed = None

    # ---- reference API ------------------------------------------------------------------------------------------
    def set_numframe(self, num_frames):
        self.num_frames = num_frames

    def set_precision(self, precision: str) -> "FSTextTransformer":
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if precision != self.precision:
            self.precision, self._packed = precision, None
        return self

    def enable_xformers_memory_efficient_attention(self):
        return None

    def set_use_memory_efficient_attention_xformers(self, valid: bool = True):
        return None

    def set_attention_slice(self, slice_size=None):
        return None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._packed = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    @property
    def device(self) -> torch.device:
        return self.learnable_query.device

    @property
    def dtype(self) -> torch.dtype:
        return self.learnable_query.dtype

    # ---- packing ----------------------------------------------------------------------------------------------------
    def _P(self, key: str) -> torch.Tensor:
        mod: nn.Module = self
        parts = key.split(".")
        for p in parts[:-1]:
            mod = mod._modules[p]
        t = mod._parameters.get(parts[-1])
        if t is None:
            t = mod._buffers[parts[-1]]
        return t.detach()

    @torch.no_grad()
    def _pack(self) -> dict:
        if self.device.type != "cuda" or self.dtype != torch.float32:
            raise RuntimeError("FSTextTransformer (seer_b200) needs fp32 parameters on a CUDA device: call .cuda() first")
        P = self._P
        f32 = lambda k: P(k).float().contiguous()
        fp32 = self.precision == "fp32"
        lin = (lambda w: packing.pack_linear(packing.split3_weight(w))) if fp32 else packing.pack_linear
        layers = []
        for n in range(self.num_layers):
            blocks = []
            for j in (0, 1):
                b = f"trf_blocks.{n}.transformer_blocks.{j}."
                t = dict(temporal=bool(j))
                for i in (1, 3) if j else (1, 2, 3):
                    t[f"ln{i}_g"], t[f"ln{i}_b"] = f32(b + f"norm{i}.weight"), f32(b + f"norm{i}.bias")
                t["qkv_w"] = lin(torch.cat([P(b + "attn1.to_q.weight"), P(b + "attn1.to_k.weight"), P(b + "attn1.to_v.weight")], 0))
                t["o1_w"], t["o1_b"] = lin(P(b + "attn1.to_out.0.weight")), f32(b + "attn1.to_out.0.bias")
                if fp32:
                    t["ff1_w"], t["ff1_b"] = lin(P(b + "ff.net.0.proj.weight")), f32(b + "ff.net.0.proj.bias")
                else:
                    t["ff1_w"], t["ff1_b"] = packing.pack_geglu(P(b + "ff.net.0.proj.weight"), P(b + "ff.net.0.proj.bias"))
                t["ff2_w"], t["ff2_b"] = lin(P(b + "ff.net.2.weight")), f32(b + "ff.net.2.bias")
                if j:
                    t["freqs"] = f32(b + "attn1.rotary_emb.freqs")
                else:
                    t["q2_w"] = lin(P(b + "attn2.to_q.weight"))
                    t["kv2_w"] = lin(torch.cat([P(b + "attn2.to_k.weight"), P(b + "attn2.to_v.weight")], 0))
                    t["o2_w"], t["o2_b"] = lin(P(b + "attn2.to_out.0.weight")), f32(b + "attn2.to_out.0.bias")
                blocks.append(t)
            layers.append(blocks)
        self._packed = dict(layers=layers, norm_g=f32("norm.weight"), norm_b=f32("norm.bias"))
        return self._packed

    def _query_tokens(self, b: int, l: int) -> torch.Tensor:
        """learnable_query + pos_embed[:, :, :l] (nearest-resized along the frame axis when num_frames differs from the
        table, unet_3d_condition.py:474-479) -> fp32 [(b f l), C].  Input-independent: parameter preprocessing."""
        pos = self.pos_embed.detach()[:, :, :l, :].float()
        if pos.shape[1] != self.num_frames:
            pos = F.interpolate(pos.permute(0, 3, 1, 2), size=(self.num_frames, l)).permute(0, 2, 3, 1)
        x = self.learnable_query.detach().float() + pos                      # (1, F, l, C)
        return x.expand(b, -1, -1, -1).reshape(b * self.num_frames * l, self.channels).contiguous()

    # ---- forward ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, context: torch.Tensor) -> torch.Tensor:
        if context.dim() != 3 or context.shape[2] != self.cross_attention_dim:
            raise ValueError(f"context must be (b, L, {self.cross_attention_dim}), got {tuple(context.shape)}")
        pk = self._packed or self._pack()
        dev = self.device
        fp32 = self.precision == "fp32"
        b, l, _ = context.shape
        Fr, C, heads = self.num_frames, self.channels, self.heads
        d = C // heads
        act = torch.float32 if fp32 else torch.bfloat16
        ctx = context.to(device=dev, dtype=torch.float32).reshape(b * l, -1).contiguous()
        ctx_op = ops.split3(ctx) if fp32 else ops.cast_bf16(ctx)

        def norm(x, g, bb):
            return ops.layernorm_f32(x, g, bb) if fp32 else ops.layernorm(x, g, bb)

        def lin(a, w, bias=None, residual=None, to_act=False):
            """a: normalised / attention activations (act dtype).  to_act: result feeds attention (act dtype), else fp32."""
            if fp32:
                return ops.gemm_ex(ops.split3(a), w, bias=bias, residual=residual).out
            return ops.gemm_ex(a, w, bias=bias, residual=residual, out_dtype=act if to_act else torch.float32).out

        def ff(t, tok):
            n3 = norm(tok, t["ln3_g"], t["ln3_b"])
            if fp32:
                hid = ops.geglu_f32(lin(n3, t["ff1_w"], t["ff1_b"]))
            else:
                hid = ops.gemm_ex(n3, t["ff1_w"], bias=t["ff1_b"], geglu=True).out
            return lin(hid, t["ff2_w"], t["ff2_b"], residual=tok)

        tok = self._query_tokens(b, l)
        for blk0, blk1 in pk["layers"]:
            # block 0: per-frame text self-attention, cross-attention to the CLIP tokens, FF (attention.py:398-427)
            qkv = lin(norm(tok, blk0["ln1_g"], blk0["ln1_b"]), blk0["qkv_w"], to_act=True)
            att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SPATIAL, heads=heads, n_outer=b * Fr,
                                Lq=l, Lk=l)
            tok = lin(att, blk0["o1_w"], blk0["o1_b"], residual=tok)
            q2 = lin(norm(tok, blk0["ln2_g"], blk0["ln2_b"]), blk0["q2_w"], to_act=True)
            kv = ops.gemm_ex(ctx_op, blk0["kv2_w"], out_dtype=act).out
            att2 = ops.attention(q2, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=b, Lq=Fr * l, Lk=l)
            tok = lin(att2, blk0["o2_w"], blk0["o2_b"], residual=tok)
            tok = ff(blk0, tok)
            # block 1: causal RoPE attention along the frame axis per (clip, token), FF (attention.py:388-396)
            qkv = lin(norm(tok, blk1["ln1_g"], blk1["ln1_b"]), blk1["qkv_w"], to_act=True)
            ops.rope_ex(qkv, l, Fr, heads, d, 0, C, blk1["freqs"])
            att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_FRAME, heads=heads, n_outer=b, F=Fr, H=l)
            tok = lin(att, blk1["o1_w"], blk1["o1_b"], residual=tok)
            tok = ff(blk1, tok)
        out = ops.layernorm_f32(tok, pk["norm_g"], pk["norm_b"])
        return out.view(b, Fr, l, C)
