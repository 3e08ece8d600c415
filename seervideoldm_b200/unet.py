"""SeerUNet — host-side mirror of the reference's module API over the seer_b200 kernels.

Drop-in for /root/reference/seer/models/unet_3d_condition.py:61-376 on the denoising hot path:
same constructor arguments, same `forward(sample, timestep, context, cond_frame=0, return_attn=False)`
signature (plus the `encoder_hidden_states=` alias), same state-dict keys/shapes
(Appendix C, loads reference checkpoints with strict=True), same exceptions style.  Underneath, the
whole forward runs on hand-written sm_100a kernels in one canonical channels-last layout
[(b f h w), C]; there is no PyTorch/CPU fallback — a missing libseer_b200.so raises.

Data layout in HBM (per evaluation, B = UNet batch after CFG):
  * residual stream   bf16 [B*F*h*w, C]            (block outputs / skip connections; fp32 with residual_stream = "fp32")
  * GEMM/conv operands bf16, same token-major shape (normalised activations, q/k/v, FF hidden)
  * packed weights    bf16 [N, K] K-major           (conv3x3: K = [Cin/64][ky][kx][64] (+ fused 1x1 shortcut))
  * text K/V          bf16 [B*F*77, 2C] per cross-attention layer, cached across the 31 DDIM steps
"""
from __future__ import annotations

import os

import math
from dataclasses import asdict
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops, packing
from .config import UNetConfig
from .weights import unet_schema


class _Node(nn.Module):
    """Container whose children are registered under the reference's attribute names (numeric names index like
    nn.ModuleList), so state_dict keys match the reference exactly."""

    def __getitem__(self, i):
        return self._modules[str(i)]

    def __len__(self):
        return len(self._modules)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("seer_b200 sub-modules are parameter containers; call SeerUNet.forward")


class _Config(SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)



class SeerUNet(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True,
                 freq_shift=0,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
                 mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280,
                 attention_head_dim=8):
        super().__init__()
        if len(block_out_channels) != 4:
            raise ValueError("SeerUNet hard-codes 4 resolution levels (unet_3d_condition.py:90-91)")
        if norm_num_groups != 32:
            raise ValueError("block GroupNorms are fixed at 32 groups in the reference (SURVEY F15)")
        if act_fn not in ("silu", "swish"):
            raise ValueError(f"unsupported act_fn {act_fn}")
        self.cfg = UNetConfig(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                              center_input_sample=center_input_sample, flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift,
                              block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                              mid_block_scale_factor=mid_block_scale_factor, act_fn=act_fn, norm_num_groups=norm_num_groups,
                              norm_eps=norm_eps, cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim)
        self.config = _Config(**asdict(self.cfg))
        self.sample_size = sample_size
        self._packed: Optional[dict] = None
        self._packed32: Optional[dict] = None
        self.rope_fuse_min_channels = 640
        self.conv_out_tensor_core = True
        # conv_in as im2col + tcgen05 GEMM (ops.conv_in_im2col): 378 -> 172 us per 16-sample evaluation, but rounding the latent to
        # bf16 moves the step error 9.7e-3 -> 1.01e-2 for 0.15 % of the step: off by default
        self.conv_in_tensor_core = False
        # dtype of the residual stream BETWEEN blocks on the bf16 path.  "bf16": every block output is stored once, as bf16, next
        # to the GroupNorm column sums of its fp32 values (what the reference's fp16 autocast does: conv / Linear outputs and
        # the residual adds are half precision there, resnet.py:206, attention.py:143); "fp32": fp32 block outputs (round 1).
        # Measured: step eps rel-L2 7.7e-3 -> 9.5e-3, 31-evaluation latents 6.9e-3 -> 9.6e-3 (budgets 2e-2 / 5e-2), -2.7 % time.
        self.residual_stream = os.environ.get("SEER_RESIDUAL_STREAM", "bf16")
        self._weights_version = 0       # bumped whenever the packed weights are dropped (captured CUDA graphs check it)
        self._kv_key = None
        self._kv: List[torch.Tensor] = []
        self.precision = "bf16"
        self._build_tree()
        self.reset_parameters()

    # ------------------------------------------------------------------ parameters / state dict
    def _build_tree(self) -> None:
        for key, shape in unet_schema(self.cfg).items():
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
                mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape), requires_grad=False))

    @torch.no_grad()
    def reset_parameters(self) -> None:
        """PyTorch-default-scale init; `proj_out` zero-initialised like the reference (attention.py:126-127)."""
        params = dict(self.named_parameters())
        for name, p in params.items():
            owner = name.split(".")[-2]
            if owner.startswith("norm") or owner == "conv_norm_out":
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            elif owner == "proj_out":
                p.zero_()
            else:
                ref = p if name.endswith("weight") else params[name[: -len("bias")] + "weight"]
                bound = 1.0 / math.sqrt(max(1, ref[0].numel()))
                p.uniform_(-bound, bound)
        self._invalidate_packed()

    def _invalidate_packed(self) -> None:
        """The parameters changed (load / init / device move): drop the packed bf16 copies and the text K/V, and bump the
        version so that CUDA graphs captured over the old packed tensors (graph.GraphedUNet) are never replayed."""
        self._packed = self._packed32 = None
        self._kv_key = None
        self._weights_version = getattr(self, "_weights_version", 0) + 1

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._invalidate_packed()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate_packed()
        return out

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
        """`SeerUNet.from_pretrained(sd15_dir, subfolder="unet", revision=..., low_cpu_mem_usage=False)` as the
        reference's scripts call it (inference.py:82-87): local `config.json` + `diffusion_pytorch_model.{bin,safetensors}`;
        a 2-D Stable Diffusion UNet inflates (its keys are a subset of the schema, the temporal attentions keep their
        initialisation).  See checkpoint.py."""
        from .checkpoint import unet_from_pretrained
        return unet_from_pretrained(cls, pretrained_model_name_or_path, subfolder=subfolder, **kwargs)

    def save_pretrained(self, save_directory, safe_serialization: bool = False, **kwargs):
        from .checkpoint import unet_save_pretrained
        unet_save_pretrained(self, save_directory, safe_serialization=safe_serialization)

    @property
    def dtype(self) -> torch.dtype:
        return self.conv_in.weight.dtype

    @property
    def device(self) -> torch.device:
        return self.conv_in.weight.device

    def set_precision(self, precision: str) -> "SeerUNet":
        """"bf16" (default, the product path: bf16 tensor-core operands and activations, fp32 accumulation / statistics; what
        the reference computes under accelerate's mixed precision) or "fp32" (parity mode, unet_fp32.py: error-compensated
        bf16 operand pairs on the same tcgen05 kernels, everything else fp32; matches the reference's fp32 forward to
        rel-L2 <= 1e-4)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self._kv_key = None
        return self

    # API parity no-ops (xformers / slicing are memory-saving switches of the reference's PyTorch path)
    def enable_xformers_memory_efficient_attention(self):
        return None

    def set_use_memory_efficient_attention_xformers(self, valid: bool = True):
        return None

    def set_attention_slice(self, slice_size=None):
        return None

    # ------------------------------------------------------------------ weight packing
    def _P(self, key: str) -> torch.Tensor:
        mod: nn.Module = self
        parts = key.split(".")
        for p in parts[:-1]:
            mod = mod._modules[p]
        t = mod._parameters.get(parts[-1])
        if t is None:
            t = mod._buffers[parts[-1]]
        return t.detach()

    def _has(self, key: str) -> bool:
        try:
            self._P(key)
            return True
        except KeyError:
            return False

    @torch.no_grad()
    def _pack(self) -> dict:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("SeerUNet (seer_b200) runs on CUDA only: move the module with .cuda() before calling it")
        if self.dtype != torch.float32:
            raise RuntimeError("keep parameters in fp32 (the reference cannot run otherwise either — SURVEY F10); "
                               "the kernels compute in bf16 with fp32 accumulation")
        P = self._P
        f32 = lambda k: P(k).float().contiguous()
        pk: dict = {}
        pk["conv_in_w"] = P("conv_in.weight").float().reshape(self.cfg.block_out_channels[0], -1).contiguous()
        pk["conv_in_b"] = f32("conv_in.bias")
        if self.cfg.in_channels == 4:               # conv_in as one 64-wide k-block of the tcgen05 GEMM (36 real taps)
            w64 = torch.zeros((pk["conv_in_w"].shape[0], 64), dtype=torch.float32, device=pk["conv_in_w"].device)
            w64[:, :36] = pk["conv_in_w"]
            pk["conv_in_w16"] = w64.to(torch.bfloat16).contiguous()
        pk["te1_w"], pk["te1_b"] = f32("time_embedding.linear_1.weight"), f32("time_embedding.linear_1.bias")
        pk["te2_w"], pk["te2_b"] = f32("time_embedding.linear_2.weight"), f32("time_embedding.linear_2.bias")
        pk["gno_g"], pk["gno_b"] = f32("conv_norm_out.weight"), f32("conv_norm_out.bias")
        pk["conv_out_w"] = packing.pack_conv_out(P("conv_out.weight"))
        pk["conv_out_b"] = f32("conv_out.bias")
        # conv_out as a tensor-core implicit GEMM: output channels zero-padded to one 64-column tile (bf16 operands, fp32 out)
        w_out = P("conv_out.weight")
        if w_out.shape[0] <= 64 and w_out.shape[1] % 64 == 0:
            w64 = torch.zeros((64,) + tuple(w_out.shape[1:]), device=dev, dtype=torch.float32)
            w64[: w_out.shape[0]] = w_out
            b64 = torch.zeros(64, device=dev, dtype=torch.float32)
            b64[: w_out.shape[0]] = P("conv_out.bias")
            pk["conv_out_w16"], pk["conv_out_b64"] = packing.pack_conv3x3(w64), b64

        temb_w, temb_b = [], []
        off = 0

        def resnet(prefix: str) -> dict:
            nonlocal off
            w1 = P(prefix + "conv1.weight")
            cout, cin = w1.shape[:2]
            r = dict(cin=cin, cout=cout, off=off, sc=self._has(prefix + "conv_shortcut.weight"))
            r["g1"], r["b1"] = f32(prefix + "norm1.weight"), f32(prefix + "norm1.bias")
            r["g2"], r["b2"] = f32(prefix + "norm2.weight"), f32(prefix + "norm2.bias")
            r["w1"] = packing.pack_conv3x3(w1)
            temb_w.append(f32(prefix + "time_emb_proj.weight"))
            temb_b.append(f32(prefix + "time_emb_proj.bias") + f32(prefix + "conv1.bias"))   # conv1 bias folded in
            off += cout
            if r["sc"]:
                r["w2"] = packing.pack_conv3x3(P(prefix + "conv2.weight"), P(prefix + "conv_shortcut.weight"))
                r["bias2"] = f32(prefix + "conv2.bias") + f32(prefix + "conv_shortcut.bias")
            else:
                r["w2"] = packing.pack_conv3x3(P(prefix + "conv2.weight"))
                r["bias2"] = f32(prefix + "conv2.bias")
            return r

        def xf(prefix: str, temporal: bool) -> dict:
            b = prefix + "transformer_blocks.0."
            t = dict(temporal=temporal, C=P(prefix + "proj_in.weight").shape[0])
            t["gn_g"], t["gn_b"] = f32(prefix + "norm.weight"), f32(prefix + "norm.bias")
            t["pin_w"], t["pin_b"] = packing.pack_conv1x1(P(prefix + "proj_in.weight")), f32(prefix + "proj_in.bias")
            t["pout_w"], t["pout_b"] = packing.pack_conv1x1(P(prefix + "proj_out.weight")), f32(prefix + "proj_out.bias")
            # LayerNorms are folded into the GEMM that consumes them (packing.fold_layernorm): weights carry gamma,
            # the bias carries beta @ W.T and the epilogue applies the per-row mean / rstd (SeerGemmDesc, LN fold).
            wqkv = torch.cat([P(b + "attn1.to_q.weight"), P(b + "attn1.to_k.weight"), P(b + "attn1.to_v.weight")], 0)
            t["qkv_w"], t["qkv_cs"], t["qkv_b"] = packing.fold_layernorm(wqkv, None, P(b + "norm1.weight"), P(b + "norm1.bias"))
            t["o1_w"], t["o1_b"] = packing.pack_linear(P(b + "attn1.to_out.0.weight")), f32(b + "attn1.to_out.0.bias")
            t["ff1_w"], t["ff1_cs"], t["ff1_b"] = packing.fold_layernorm(P(b + "ff.net.0.proj.weight"), P(b + "ff.net.0.proj.bias"),
                                                                         P(b + "norm3.weight"), P(b + "norm3.bias"), geglu=True)
            t["ff2_w"], t["ff2_b"] = packing.pack_linear(P(b + "ff.net.2.weight")), f32(b + "ff.net.2.bias")
            if temporal:
                t["freqs"] = f32(b + "attn1.rotary_emb.freqs")
            else:
                t["q2_w"], t["q2_cs"], t["q2_b"] = packing.fold_layernorm(P(b + "attn2.to_q.weight"), None, P(b + "norm2.weight"),
                                                                          P(b + "norm2.bias"))
                t["kv2_w"] = packing.pack_kv(P(b + "attn2.to_k.weight"), P(b + "attn2.to_v.weight"))
                t["o2_w"], t["o2_b"] = packing.pack_linear(P(b + "attn2.to_out.0.weight")), f32(b + "attn2.to_out.0.bias")
            return t

        n = len(self.cfg.block_out_channels)
        L = self.cfg.layers_per_block
        pk["down"] = []
        for i in range(n):
            blk = dict(res=[], attn=[], tattn=[], down=None)
            for j in range(L):
                blk["res"].append(resnet(f"down_blocks.{i}.resnets.{j}."))
                if i < n - 1:
                    blk["attn"].append(xf(f"down_blocks.{i}.attentions.{j}.", False))
                    blk["tattn"].append(xf(f"down_blocks.{i}.temporal_attentions.{j}.", True))
            if i < n - 1:
                blk["down"] = (packing.pack_conv3x3(P(f"down_blocks.{i}.downsamplers.0.conv.weight")),
                               f32(f"down_blocks.{i}.downsamplers.0.conv.bias"))
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
                wu = P(f"up_blocks.{i}.upsamplers.0.conv.weight")
                blk["up"] = (packing.pack_conv3x3(wu), f32(f"up_blocks.{i}.upsamplers.0.conv.bias"), packing.pack_upsample_phases(wu))
            pk["up"].append(blk)
        pk["temb_w"] = torch.cat(temb_w, 0).contiguous()
        pk["temb_b"] = torch.cat(temb_b, 0).contiguous()
        self._packed = pk
        return pk

    # ------------------------------------------------------------------ operators
    @staticmethod
    def _emit(g: "ops.GemmOut", want: str, s16: bool = False):
        """Block output record (fp32 stream or None, GroupNorm col_stats, bf16 copy or None) of the GEMM that produced it.
        `want`: "f32" (stream tensor + statistics, the residual stream between blocks), "both" (the stride-2 Downsample3D conv
        also reads it: + bf16 copy on the fp32 stream), "bf16" (bf16 only, no statistics: the block feeds an Upsample3D conv and
        nothing else).  `s16`: the stream itself is bf16 — the record carries no fp32 tensor."""
        if want == "bf16":
            return (None, None, g.out)
        if s16:
            return (None, g.col_stats, g.out)
        return (g.out, g.col_stats, g.out16)

    @staticmethod
    def _stream(x):
        """The tensor of a stream record that GroupNorm / the residual add read: fp32 when present, else the bf16 one."""
        return x[0] if x[0] is not None else x[2]

    def _resnet(self, r: dict, x1, x2, B, F, H, W, temb_all, want: str = "f32"):
        """ResnetBlock3D.forward (resnet.py:174-208) on the virtual concat [x1 | x2].  Activations travel as
        (fp32 tensor, col_stats, bf16 copy or None) records: the GEMM that produced a tensor also emitted the per-channel
        partial sums the next GroupNorm needs, so no statistics pass re-reads the activation.  conv1's output is stored as
        bf16 only (its single reader is GroupNorm 2; the statistics are those of the stored bf16 values)."""
        T = F * H * W
        eps = self.cfg.norm_eps
        cin, cout = r["cin"], r["cout"]
        s16 = x1[0] is None                        # bf16 residual stream: records carry (None, col_stats, bf16)
        t1, s1 = self._stream(x1), x1[1]
        t2, s2 = (self._stream(x2), x2[1]) if x2 is not None else (None, None)
        if r["sc"] and s16 and t2 is None:
            # bf16 stream, no concat: the block input already is the bf16 operand of the fused 1x1 shortcut
            h, raw = ops.groupnorm(t1, None, B, r["g1"], r["b1"], eps, True, stats1=s1), t1
        elif r["sc"]:
            h, raw = ops.groupnorm(t1, t2, B, r["g1"], r["b1"], eps, True, want_raw=True, stats1=s1, stats2=s2)
        else:
            if t2 is not None:
                raise RuntimeError("concat input without a shortcut conv cannot occur in this architecture")
            h, raw = ops.groupnorm(t1, None, B, r["g1"], r["b1"], eps, True, stats1=s1), None
        tb = temb_all[:, r["off"]: r["off"] + cout]
        if T % 32 == 0:
            c1 = ops.conv3x3_ex(h.view(B * F, H, W, cin), r["w1"], bias=tb, bias_div=T, col_stats=True, out_dtype=torch.bfloat16)
        else:       # no producer statistics for ragged samples: GroupNorm 2 takes its own pass over an fp32 tensor
            c1 = ops.conv3x3_ex(h.view(B * F, H, W, cin), r["w1"], bias=tb, bias_div=T)
        h2 = ops.groupnorm(c1.out, None, B, r["g2"], r["b2"], eps, True, stats1=c1.col_stats)
        o32 = want != "bf16" and not s16
        kw = dict(bias=r["bias2"], col_stats=(want != "bf16"), out_dtype=torch.float32 if o32 else torch.bfloat16,
                  also_bf16=(want == "both" and o32))
        if r["sc"]:
            c2 = ops.conv3x3_ex(h2.view(B * F, H, W, cout), r["w2"], a2=raw, **kw)
        else:
            c2 = ops.conv3x3_ex(h2.view(B * F, H, W, cout), r["w2"], residual=t1, **kw)
        return self._emit(c2, want, s16)

    def _ff(self, t: dict, tok, rstats, out_rows=None):
        """x + FF(LN3(x)) -> bf16 (feeds proj_out only).  attention.py:244,323 + 744-747,791-793.  LN3 is folded into
        the GEGLU projection, which reads the raw bf16 token stream."""
        hid = ops.gemm_ex(tok, t["ff1_w"], bias=t["ff1_b"], geglu=True, ln=(rstats, t["ff1_cs"], 1e-5)).out
        return ops.gemm_ex(hid, t["ff2_w"], bias=t["ff2_b"], residual=tok, out=out_rows, out_dtype=torch.bfloat16).out

    def _transformer(self, t: dict, x, B, F, H, W, kv, cond_frame, want: str = "f32", dup: bool = False):
        """SpatialTransformer3D.forward (attention.py:129-145) with its text (:308-327) or temporal (:231-248) block.
        The token stream inside the block is bf16 (what the reference computes under autocast: Linear outputs and the
        residual adds are low precision there too) with per-row (sum, sumsq) written by the producing GEMM's epilogue for
        the LayerNorm folded into the next projection; the block input / output follow `residual_stream` (bf16 by default).

        `dup` (text block only): `x` holds the FIRST HALF of a CFG batch whose two halves are identical up to here (same
        latents, same timestep — ddim_video.py:199-203); everything that does not see the text context (GroupNorm, proj_in,
        LN1 + q/k/v, self-attention, to_out, LN2 + cross-attention query) is computed once on B/2 samples, then the
        token stream, the query and the block input are duplicated and the block continues on B samples."""
        C, heads = t["C"], self.cfg.heads
        d = C // heads
        hw, T = H * W, F * H * W
        bf = torch.bfloat16
        s16 = x[0] is None
        xt, xs = self._stream(x), x[1]
        Bfull = B
        if dup:
            if t["temporal"] or B % 2:
                raise RuntimeError("dup: text block on an even CFG batch only")
            B = B // 2
        M = B * T
        hn = ops.groupnorm(xt, None, B, t["gn_g"], t["gn_b"], 1e-6, False, stats1=xs)
        r = ops.gemm_ex(hn, t["pin_w"], bias=t["pin_b"], out_dtype=bf, row_stats=True)
        rope = tab = None
        if t["temporal"] and t["freqs"].numel() == 16:
            # fp16 (cos, sin) table of this clip length, built once per layer (outside any graph capture: the warm-up
            # evaluations fill the cache); read by the fused epilogue below or by the vectorised stand-alone pass
            tab = t.setdefault("rope_tabs", {}).get(T)
            if tab is None:
                tab = t["rope_tabs"][T] = ops.rope_table(t["freqs"], T)
        if tab is not None and T % 32 == 0 and C >= self.rope_fuse_min_channels:
            # rotary embedding of q and k (attention.py:649-651) fused into the projection's epilogue.
            # Only where the GEMM's main loop hides the extra epilogue work (K = C >= 640): at the 320-channel level the
            # projection is epilogue-bound and the fused form measured SLOWER than the separate pass (+119 us vs 94 us,
            # profiles/r2_gemm_probe.txt), so level 0 keeps `rope_inplace`.
            rope = (tab, 2 * C, d)
        qkv = ops.gemm_ex(r.out, t["qkv_w"], bias=t["qkv_b"], out_dtype=bf, ln=(r.row_stats, t["qkv_cs"], 1e-5), rope=rope).out   # [M, 3C]
        if t["temporal"]:
            if rope is None:
                ops.rope_inplace(qkv, T, heads, d, 0, C, t["freqs"], tab=tab)
            att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B,
                                F=F, H=H, W=W)
        else:
            att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SPATIAL, heads=heads,
                                n_outer=B * F, Lq=hw, Lk=hw)
        r = ops.gemm_ex(att, t["o1_w"], bias=t["o1_b"], residual=r.out, out_dtype=bf, row_stats=True)
        if not t["temporal"]:
            q2 = ops.gemm_ex(r.out, t["q2_w"], bias=t["q2_b"], out_dtype=bf, ln=(r.row_stats, t["q2_cs"], 1e-5)).out
            if dup:
                # the two CFG branches diverge here: same queries / token stream / block input, different text K/V
                q2 = torch.cat([q2, q2])
                r = ops.GemmOut(torch.cat([r.out, r.out]))
                xt = torch.cat([xt, xt])
                B, M = Bfull, Bfull * T
            Lk = kv.shape[0] // (B * F)
            att2 = ops.attention(q2, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=B * F, Lq=hw, Lk=Lk)
            r = ops.gemm_ex(att2, t["o2_w"], bias=t["o2_b"], residual=r.out, out_dtype=bf, row_stats=True)
        tok, rstats = r.out, r.row_stats
        if t["temporal"] and cond_frame > 0:
            # the first cond_frame frames of every clip bypass the feed-forward (attention.py:240-246): one launch pair per sample.
            # (Running the feed-forward over ALL tokens in one launch and copying the bypassed rows back is bit-identical for the
            #  rows that keep it, and was measured at the bench shape: 3.693 vs 3.701 clips/s — the per-sample launches win.)
            c0 = min(cond_frame, F) * hw
            y = torch.empty((M, C), device=xt.device, dtype=bf)
            # two strided copies for the whole batch instead of two per sample (512 tiny launches per evaluation at batch 16; same-box
            # A/B: device-resident throughput unchanged, end to end 3.61 -> 3.69 clips/s):
            # the bypassed rows, and each sample's LayerNorm row sums of the rows that keep their feed-forward, made contiguous
            y.view(B, T, C)[:, :c0].copy_(tok.view(B, T, C)[:, :c0])
            if c0 < T:
                parts = rstats.shape[0]
                rs_all = rstats.view(parts, B, T, 2)[:, :, c0:].permute(1, 0, 2, 3).contiguous()      # [B][parts][T - c0][2]
                for b in range(B):
                    self._ff(t, tok[b * T + c0:(b + 1) * T], rs_all[b], out_rows=y[b * T + c0:(b + 1) * T])
        else:
            y = self._ff(t, tok, rstats)
        o32 = want != "bf16" and not s16
        o = ops.gemm_ex(y, t["pout_w"], bias=t["pout_b"], residual=xt, col_stats=(want != "bf16"), also_bf16=(want == "both" and o32),
                        out_dtype=torch.float32 if o32 else bf)
        return self._emit(o, want, s16)

    def _cross_layers(self, pk: dict) -> List[dict]:
        return [a for blk in pk["down"] for a in blk["attn"]] + [pk["mid"]["attn"]] + [a for blk in pk["up"] for a in blk["attn"]]

    def compute_context_kv(self, context: torch.Tensor, out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """K/V projections of the text context for every cross-attention layer: [B*F*L, 2C] bf16 each
        (attention.py:517-518 with the per-frame context of :314-315).  `out` recomputes into existing buffers."""
        if self.device.type == "cuda" and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.compute_context_kv(context, out)
        if self.precision == "fp32":
            from . import unet_fp32
            if self._packed32 is None:
                self._packed32 = unet_fp32.pack_fp32(self)
            return unet_fp32.context_kv(self, self._packed32, context, out)
        pk = self._packed or self._pack()
        ctx = ops.cast_bf16(context.reshape(-1, context.shape[-1]).float().contiguous())
        layers = self._cross_layers(pk)
        if out is None:
            return [ops.gemm(ctx, a["kv2_w"], out_dtype=torch.bfloat16) for a in layers]
        for a, o in zip(layers, out):
            ops.gemm(ctx, a["kv2_w"], out=o)
        return out

    def _context_kv(self, pk: dict, context: torch.Tensor) -> List[torch.Tensor]:
        """K/V of every cross-attention layer depend only on the text context -> computed once per clip and reused
        for all 31 DDIM evaluations (SURVEY §7.1 'exploitable redundancy' (i)).  The cache is keyed on tensor identity
        + version and holds a reference: a data_ptr()-only key would go stale when the caching allocator hands a
        freed context's address to a new tensor."""
        if (self._kv_key is not None and self._kv_key[0] is context and self._kv_key[1] == context._version
                and self._kv_key[2] == self.precision):
            return self._kv
        self._kv = self.compute_context_kv(context)
        self._kv_key = (context, context._version, self.precision)
        return self._kv

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep, context: Optional[torch.Tensor] = None, cond_frame: int = 0,
                return_attn: bool = False, encoder_hidden_states: Optional[torch.Tensor] = None,
                cfg_shared_input: bool = False) -> torch.Tensor:
        """`cfg_shared_input=True` (not in the reference's signature; DDIMSampler sets it) promises that the batch is the
        sampler's CFG batch `[x; x]` with one timestep — identical halves that differ only in `context` — so the layers in
        front of the first cross-attention are evaluated once instead of twice (bit-identical result)."""
        self._cfg_shared = bool(cfg_shared_input)
        # every kernel launches on the current stream of the CURRENT device and tensor maps are encoded in its context:
        # make the model's device current for the whole evaluation (a module on cuda:1 called while cuda:0 is current)
        if self.device.type == "cuda":
            with torch.cuda.device(self.device):
                return self._forward(sample, timestep, context, cond_frame, return_attn, encoder_hidden_states)
        return self._forward(sample, timestep, context, cond_frame, return_attn, encoder_hidden_states)

    def _forward(self, sample, timestep, context, cond_frame, return_attn, encoder_hidden_states) -> torch.Tensor:
        if context is None:
            context = encoder_hidden_states
        if context is None:
            raise ValueError("context (B, F, L, cross_attention_dim) is required")
        if return_attn:
            raise NotImplementedError("return_attn=True (pre-softmax score dump) is not on the sampling path")
        if sample.dim() != 5 or context.dim() != 4:
            raise ValueError("expected sample (B,C,F,H,W) and context (B,F,L,D)")
        if self.precision == "fp32":
            if self.device.type != "cuda" or self.dtype != torch.float32:
                raise RuntimeError("SeerUNet (seer_b200) needs fp32 parameters on a CUDA device")
            from . import unet_fp32
            if self._packed32 is None:
                self._packed32 = unet_fp32.pack_fp32(self)
            pk = self._packed32
        else:
            pk = self._packed or self._pack()
        cfg = self.cfg
        dev = self.device
        B, Cin, F, H, W = sample.shape
        if Cin != cfg.in_channels or context.shape[0] != B or context.shape[1] != F or context.shape[3] != cfg.cross_attention_dim:
            raise ValueError(f"shape mismatch: sample {tuple(sample.shape)} context {tuple(context.shape)}")
        if H % 8 or W % 8:
            raise ValueError("latent height/width must be multiples of 8 (three stride-2 levels)")
        sample = sample.to(device=dev, dtype=torch.float32)
        if cfg.center_input_sample:
            sample = 2 * sample - 1.0
        # 1. time (unet_3d_condition.py:298-308)
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=dev)
        elif t.dim() == 0:
            t = t[None].to(dev)
        t = t.to(dev).broadcast_to((B,)).to(torch.float32).contiguous()
        if self.precision == "fp32":
            return unet_fp32.forward(self, pk, sample, t, context, cond_frame, self._context_kv(pk, context.to(dev)))
        temb = ops.timestep_embedding(t, cfg.block_out_channels[0], float(cfg.freq_shift), cfg.flip_sin_to_cos)
        e1 = ops.small_linear(temb, pk["te1_w"], pk["te1_b"], silu_out=True)
        emb = ops.small_linear(e1, pk["te2_w"], pk["te2_b"], silu_out=True)   # = SiLU(emb): emb is consumed only through time_emb_proj(SiLU(emb)), resnet.py:190-192
        temb_all = ops.small_linear(emb, pk["temb_w"], pk["temb_b"])      # all 22 time_emb_proj at once
        kvs = iter(self._context_kv(pk, context.to(dev)))

        # 2. conv_in -> token-major residual stream
        # CFG batch with identical halves: conv_in, the first ResNet block and the context-free front of the first text block
        # run on the first half only (ddim_video.py:199-203 builds x_in = cat([x] * 2), t_in = cat([t] * 2))
        shared = (getattr(self, "_cfg_shared", False) and B % 2 == 0 and bool(pk["down"][0]["attn"]) and (F * H * W) % 32 == 0)
        Bc = B // 2 if shared else B
        # bf16 residual stream: needs producer statistics at every level (32-row slabs must not straddle samples)
        s16 = self.residual_stream == "bf16" and (F * (H // 8) * (W // 8)) % 32 == 0
        Mc = Bc * F * H * W
        if self.conv_in_tensor_core and "conv_in_w16" in pk and Mc % 32 == 0 and pk["conv_in_w16"].shape[0] % 64 == 0:
            # bf16 operands (the latent is rounded once, like every other activation operand), fp32 accumulate; the fp32 SIMT
            # kernel is FP32-pipe bound (380 us for 8 clips)
            g0 = ops.gemm_ex(ops.conv_in_im2col(sample[:Bc].contiguous()), pk["conv_in_w16"], bias=pk["conv_in_b"], col_stats=True,
                             out_dtype=torch.bfloat16 if s16 else torch.float32)
            ci, cst = g0.out, g0.col_stats
        else:
            ci, cst = ops.conv_in(sample[:Bc].contiguous(), pk["conv_in_w"], pk["conv_in_b"], col_stats=True,
                                  out_dtype=torch.bfloat16 if s16 else torch.float32)
        x = (None, cst, ci) if s16 else (ci, cst, None)                 # (fp32 stream | None, col_stats, bf16)
        h, w = H, W
        dup2 = lambda t_: None if t_ is None else torch.cat([t_, t_])
        skips: List[tuple] = [tuple(dup2(t_) for t_ in x) if shared else x]
        n = len(cfg.block_out_channels)
        # 3. down
        for i, blk in enumerate(pk["down"]):
            nres = len(blk["res"])
            for j, r in enumerate(blk["res"]):
                # the block's last tensor also feeds the stride-2 conv, which reads bf16
                last = "both" if (blk["down"] is not None and j == nres - 1) else "f32"
                first = shared and i == 0 and j == 0
                x = self._resnet(r, x, None, Bc if first else B, F, h, w, temb_all, want="f32" if blk["attn"] else last)
                if blk["attn"]:
                    x = self._transformer(blk["attn"][j], x, B, F, h, w, next(kvs), cond_frame, dup=first)
                    x = self._transformer(blk["tattn"][j], x, B, F, h, w, None, cond_frame, want=last)
                skips.append(x)
            if blk["down"] is not None:
                wd, bd = blk["down"]
                # Downsample3D (resnet.py:95-104): stride-2 / pad-1 conv as an implicit GEMM over strided TMA boxes
                odt = torch.bfloat16 if s16 else torch.float32
                dn = ops.gemm_ex(None, wd, x_img=x[2].view(B * F, h, w, x[2].shape[1]), conv_stride=2, bias=bd, col_stats=True, out_dtype=odt)
                if dn is None:       # geometry outside the TMA-box tiling: explicit im2col (still seer_b200 kernels)
                    cols = ops.im2col3x3(x[2].view(B * F, h, w, x[2].shape[1]), stride=2)
                    dn = ops.gemm_ex(cols, wd, bias=bd, col_stats=True, out_dtype=odt)
                x = (None, dn.col_stats, dn.out) if s16 else (dn.out, dn.col_stats, None)
                h, w = h // 2, w // 2
                skips.append(x)
        # 4. mid
        m = pk["mid"]
        x = self._resnet(m["res"][0], x, None, B, F, h, w, temb_all)
        x = self._transformer(m["attn"], x, B, F, h, w, next(kvs), cond_frame)
        x = self._transformer(m["tattn"], x, B, F, h, w, None, cond_frame)
        x = self._resnet(m["res"][1], x, None, B, F, h, w, temb_all)
        # 5. up
        for i, blk in enumerate(pk["up"]):
            nres = len(blk["res"])
            for j, r in enumerate(blk["res"]):
                # the block's last tensor feeds only the Upsample3D conv: bf16, no statistics
                last = "bf16" if (blk["up"] is not None and j == nres - 1) else "f32"
                x = self._resnet(r, x, skips.pop(), B, F, h, w, temb_all, want="f32" if blk["attn"] else last)
                if blk["attn"]:
                    x = self._transformer(blk["attn"][j], x, B, F, h, w, next(kvs), cond_frame)
                    x = self._transformer(blk["tattn"][j], x, B, F, h, w, None, cond_frame, want=last)
            if blk["up"] is not None:
                x = self._upsample_conv(blk["up"], x[2], B, F, h, w, s16)
                h, w = 2 * h, 2 * w
        # 6. out: GN -> SiLU -> conv_out, fp32, back to (B, C, F, H, W)
        if self.conv_out_tensor_core and "conv_out_w16" in pk and x[1] is not None:
            # bf16 operands on the tcgen05 implicit-GEMM conv (N = 64: the 4 real channels + zero rows), fp32 accumulate / out:
            # the fp32 SIMT kernel below is FP32-pipe-bound at ~0.9 ms per evaluation, this path ~0.2 ms; the operand rounding
            # adds ~1.5e-3 to the step's 7e-3 (rel-L2, in quadrature)
            y16 = ops.groupnorm(self._stream(x), None, B, pk["gno_g"], pk["gno_b"], cfg.norm_eps, True, stats1=x[1])
            r = ops.conv3x3_ex(y16.view(B * F, h, w, y16.shape[1]), pk["conv_out_w16"], bias=pk["conv_out_b64"])
            return ops.tokens_to_nchw(r.out, B, cfg.out_channels, F, h, w)
        xs_ = self._stream(x)
        y = ops.groupnorm(xs_ if xs_.dtype == torch.float32 else xs_.float(), None, B, pk["gno_g"], pk["gno_b"], cfg.norm_eps, True,
                          out_dtype=torch.float32, stats1=x[1])
        return ops.conv_out(y, pk["conv_out_w"], pk["conv_out_b"], B, F, h, w)

    def _upsample_conv(self, up: tuple, x16: torch.Tensor, B: int, F: int, h: int, w: int, s16: bool = False):
        """Upsample3D.forward (resnet.py:47-61): nearest 2x + conv3x3, computed as four 2x2-tap convs on the LOW-res image
        (packing.pack_upsample_phases) whose epilogues scatter their rows to the four pixel phases of the output — the
        upsampled tensor is never materialised and 5/9 of the FLOPs disappear."""
        wu, bu, phases = up
        C = x16.shape[1]
        n_img = B * F
        M = n_img * 4 * h * w
        img = x16.view(n_img, h, w, C)
        if (w & (w - 1)) == 0 and (F * h * w) % 32 == 0:        # 32-row statistics slabs must not straddle samples
            out = torch.empty((M, wu.shape[0]), device=x16.device, dtype=torch.bfloat16 if s16 else torch.float32)
            st = torch.empty((M // 32, wu.shape[0], 2), device=x16.device, dtype=torch.float32)
            ok = True
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                r = ops.gemm_ex(None, phases[ph], x_img=img, conv_taps=(2, 2, px - 1, py - 1), up_phase=1 + ph, bias=bu, out=out,
                                col_stats=st)
                if r is None:
                    ok = False
                    break
            if ok:
                return (None, st, out) if s16 else (out, st, None)
        # geometry outside the TMA-box tiling: materialise the upsampled image (nearest-2x of a bf16 tensor via torch indexing
        # would be a PyTorch op on the product path, so go through the fp32 upsample kernel)
        up_img = ops.upsample2x(x16.float(), n_img, h, w)
        uc = ops.conv3x3_ex(up_img, wu, bias=bu, col_stats=True, out_dtype=torch.bfloat16 if s16 else torch.float32)
        return (None, uc.col_stats, uc.out) if s16 else (uc.out, uc.col_stats, None)
