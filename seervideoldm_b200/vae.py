"""AutoencoderKL (Stable Diffusion 1.5 VAE) on the seer_b200 kernels — the network on either side of the DDIM loop
(SURVEY §8f rank 2).

The reference decodes the sampled latents with `vae.decode(1 / 0.18215 * z).sample` (utils/ddim_sampling_utils.py:37-41) and
encodes the reference frames with `vae.encode(frames).latent_dist.sample() * 0.18215` (inference.py:186-187), where `vae` is
diffusers 0.10.2's `AutoencoderKL.from_pretrained(sd15, subfolder="vae")` (inference.py:76-81).  This module is a drop-in for
that object on those two calls: same constructor configuration, same state-dict keys / shapes as the published checkpoint
(83 653 863 parameters, loads with strict=True), `decode(z).sample`, `encode(x).latent_dist.sample() / .mode() / .mean`.

Everything runs in the UNet's channels-last token layout [(n y x), C] on the same kernels:
  * 3x3 convs = tcgen05 implicit GEMM over TMA boxes (row-segment tiles for the 256-pixel-wide levels), the 1x1 ResNet
    shortcuts fused into conv2's K loop, GroupNorm statistics from the producing conv's epilogue;
  * Upsample2D (nearest 2x + conv) = four 2x2-tap phase convs on the low-res image; Downsample2D (pad right/bottom, stride 2)
    = the strided-TMA conv with tap offsets (0, 0);
  * the single-head d = 512 AttentionBlock = three tcgen05 GEMMs per image (Q K^T, V^T = W_v X^T computed transposed so no
    transpose pass exists, P V) around one row-softmax kernel; the value bias is folded into the output projection's bias
    (softmax rows sum to one);
  * the 4-channel boundaries (conv_in / conv_out, post_quant_conv / quant_conv) stay fp32 like the UNet's.
bf16 tensor-core operands with fp32 accumulation and an fp32 residual stream; no CPU / PyTorch fallback.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops, packing
from .unet import _Node

Shape = Tuple[int, ...]
SD_VAE_SCALE = 0.18215


def vae_schema(in_channels: int = 3, out_channels: int = 3, block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2,
               latent_channels: int = 4) -> "OrderedDict[str, Shape]":
    """State-dict keys / shapes of diffusers 0.10.2 `AutoencoderKL` in its registration order."""
    s: "OrderedDict[str, Shape]" = OrderedDict()

    def conv(p, cout, cin, k):
        s[p + ".weight"], s[p + ".bias"] = (cout, cin, k, k), (cout,)

    def norm(p, c):
        s[p + ".weight"], s[p + ".bias"] = (c,), (c,)

    def lin(p, cout, cin):
        s[p + ".weight"], s[p + ".bias"] = (cout, cin), (cout,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin); conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def mid(p, c):
        p_a = p + ".attentions.0"
        norm(p_a + ".group_norm", c)
        for n in ("query", "key", "value", "proj_attn"):
            lin(f"{p_a}.{n}", c, c)
        resnet(p + ".resnets.0", c, c)
        resnet(p + ".resnets.1", c, c)

    ch = list(block_out_channels)
    conv("encoder.conv_in", ch[0], in_channels, 3)
    cin = ch[0]
    for i, c in enumerate(ch):
        for j in range(layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin, c)
            cin = c
        if i < len(ch) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
    mid("encoder.mid_block", ch[-1])
    norm("encoder.conv_norm_out", ch[-1])
    conv("encoder.conv_out", 2 * latent_channels, ch[-1], 3)
    conv("decoder.conv_in", ch[-1], latent_channels, 3)
    rev = ch[::-1]
    cin = rev[0]
    for i, c in enumerate(rev):
        for j in range(layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin, c)
            cin = c
        if i < len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    mid("decoder.mid_block", rev[0])
    norm("decoder.conv_norm_out", ch[0])
    conv("decoder.conv_out", out_channels, ch[0], 3)
    conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    conv("post_quant_conv", latent_channels, latent_channels, 1)
    return s


def random_vae_state_dict(seed: int = 0, **cfg) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic VAE weights (fp32, CPU): convs / linears at PyTorch's default scale, norms 1 + N(0, 0.1) / N(0, 0.1).
    (There is no network access for the published checkpoint; parity tests use these on both sides.)"""
    out: Dict[str, torch.Tensor] = OrderedDict()
    schema = vae_schema(**cfg)
    for i, (key, shape) in enumerate(schema.items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + i)
        owner = key.split(".")[-2]
        if "norm" in owner:
            out[key] = (1.0 if key.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = shape if key.endswith("weight") else schema[key[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out


class DiagonalGaussianDistribution:
    """diffusers.models.vae.DiagonalGaussianDistribution over moments (n, 2c, h, w) = [mean | logvar]."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 3,
                 down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                 block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, act_fn: str = "silu", latent_channels: int = 4,
                 norm_num_groups: int = 32, sample_size: int = 512):
        super().__init__()
        if norm_num_groups != 32 or act_fn not in ("silu", "swish"):
            raise ValueError("seer_b200 AutoencoderKL: GroupNorm(32) + SiLU only (the SD-1.5 VAE configuration)")
        if in_channels > 4 or out_channels > 4 or latent_channels != 4:
            raise ValueError("seer_b200 AutoencoderKL: <= 4 image channels and 4 latent channels (fp32 boundary conv kernels)")
        if any(c % 64 for c in block_out_channels):
            raise ValueError("block_out_channels must be multiples of 64 (tcgen05 K blocks)")
        self.cfg = dict(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                        layers_per_block=layers_per_block, latent_channels=latent_channels)
        self.config = SimpleNamespace(**self.cfg, down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types), act_fn=act_fn,
                                      norm_num_groups=norm_num_groups, sample_size=sample_size)
        self.eps = 1e-6
        self.images_per_chunk = 32          # images decoded / encoded per pass (bounds the full-resolution activations)
        self._packed: Optional[dict] = None
        for key, shape in vae_schema(**self.cfg).items():
            parts = key.split(".")
            mod: nn.Module = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Node())
                mod = mod._modules[p]
            mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))

    # ------------------------------------------------------------------ parameters
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
        return self.post_quant_conv.weight.device

    @property
    def dtype(self) -> torch.dtype:
        return self.post_quant_conv.weight.dtype

    def _P(self, key: str) -> torch.Tensor:
        mod: nn.Module = self
        parts = key.split(".")
        for p in parts[:-1]:
            mod = mod._modules[p]
        return mod._parameters[parts[-1]].detach()

    @torch.no_grad()
    def _pack(self) -> dict:
        if self.device.type != "cuda" or self.dtype != torch.float32:
            raise RuntimeError("AutoencoderKL (seer_b200) needs fp32 parameters on a CUDA device (no CPU path)")
        P = self._P
        f32 = lambda k: P(k).float().contiguous()
        ch = list(self.cfg["block_out_channels"])
        L = self.cfg["layers_per_block"]

        def resnet(p):
            w1 = P(p + ".conv1.weight")
            cout, cin = w1.shape[:2]
            r = dict(cin=cin, cout=cout, sc=cin != cout, g1=f32(p + ".norm1.weight"), b1=f32(p + ".norm1.bias"),
                     g2=f32(p + ".norm2.weight"), b2=f32(p + ".norm2.bias"), w1=packing.pack_conv3x3(w1), bias1=f32(p + ".conv1.bias"))
            if r["sc"]:
                r["w2"] = packing.pack_conv3x3(P(p + ".conv2.weight"), P(p + ".conv_shortcut.weight"))
                r["bias2"] = f32(p + ".conv2.bias") + f32(p + ".conv_shortcut.bias")
            else:
                r["w2"] = packing.pack_conv3x3(P(p + ".conv2.weight"))
                r["bias2"] = f32(p + ".conv2.bias")
            return r

        def mid(p):
            a = p + ".attentions.0"
            wp, bv = P(a + ".proj_attn.weight").float(), f32(a + ".value.bias")
            return dict(res=[resnet(p + ".resnets.0"), resnet(p + ".resnets.1")],
                        g=f32(a + ".group_norm.weight"), b=f32(a + ".group_norm.bias"),
                        wq=packing.pack_linear(P(a + ".query.weight")), bq=f32(a + ".query.bias"),
                        wk=packing.pack_linear(P(a + ".key.weight")), bk=f32(a + ".key.bias"),
                        wv=packing.pack_linear(P(a + ".value.weight")),
                        wp=packing.pack_linear(wp), bp=(f32(a + ".proj_attn.bias") + wp @ bv).contiguous())

        pk: dict = {}
        # decoder
        pk["pq_w"] = P("post_quant_conv.weight").float().reshape(4, 4).contiguous()
        pk["pq_b"] = f32("post_quant_conv.bias")
        pk["dec_in_w"] = P("decoder.conv_in.weight").float().reshape(ch[-1], -1).contiguous()
        pk["dec_in_b"] = f32("decoder.conv_in.bias")
        pk["dec_mid"] = mid("decoder.mid_block")
        pk["dec_up"] = []
        for i in range(len(ch)):
            blk = dict(res=[resnet(f"decoder.up_blocks.{i}.resnets.{j}") for j in range(L + 1)], up=None)
            if i < len(ch) - 1:
                wu = P(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight")
                blk["up"] = (packing.pack_conv3x3(wu), f32(f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"), packing.pack_upsample_phases(wu))
            pk["dec_up"].append(blk)
        pk["dec_no_g"], pk["dec_no_b"] = f32("decoder.conv_norm_out.weight"), f32("decoder.conv_norm_out.bias")
        pk["dec_out_w"] = packing.pack_conv_out(P("decoder.conv_out.weight"))
        pk["dec_out_b"] = f32("decoder.conv_out.bias")
        # encoder (conv_in takes 4 input channels: the image is zero-padded from in_channels to 4)
        w_in = P("encoder.conv_in.weight").float()
        w_in4 = torch.zeros((w_in.shape[0], 4, 3, 3), device=w_in.device)
        w_in4[:, : w_in.shape[1]] = w_in
        pk["enc_in_w"] = w_in4.reshape(w_in.shape[0], -1).contiguous()
        pk["enc_in_b"] = f32("encoder.conv_in.bias")
        pk["enc_down"] = []
        for i in range(len(ch)):
            blk = dict(res=[resnet(f"encoder.down_blocks.{i}.resnets.{j}") for j in range(L)], down=None)
            if i < len(ch) - 1:
                blk["down"] = (packing.pack_conv3x3(P(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight")),
                               f32(f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"))
            pk["enc_down"].append(blk)
        pk["enc_mid"] = mid("encoder.mid_block")
        pk["enc_no_g"], pk["enc_no_b"] = f32("encoder.conv_norm_out.weight"), f32("encoder.conv_norm_out.bias")
        w_out = P("encoder.conv_out.weight")
        pk["enc_out_w"] = [packing.pack_conv_out(w_out[i: i + 4]) for i in range(0, w_out.shape[0], 4)]
        pk["enc_out_b"] = [f32("encoder.conv_out.bias")[i: i + 4].contiguous() for i in range(0, w_out.shape[0], 4)]
        pk["q_w"] = P("quant_conv.weight").float().reshape(w_out.shape[0], w_out.shape[0]).contiguous()
        pk["q_b"] = f32("quant_conv.bias")
        self._packed = pk
        return pk

    # ------------------------------------------------------------------ blocks (records: (fp32 tensor, col_stats, bf16 copy))
    def _resnet(self, r: dict, x, n: int, H: int, W: int, want: str = "f32"):
        """diffusers ResnetBlock2D, temb = None: x + conv2(silu(gn2(conv1(silu(gn1(x)))))) (+ 1x1 shortcut fused into conv2)."""
        t1, s1 = x[:2]
        cin, cout = r["cin"], r["cout"]
        if r["sc"]:
            h, raw = ops.groupnorm(t1, None, n, r["g1"], r["b1"], self.eps, True, want_raw=True, stats1=s1)
        else:
            h, raw = ops.groupnorm(t1, None, n, r["g1"], r["b1"], self.eps, True, stats1=s1), None
        stats_ok = (H * W) % 32 == 0
        c1 = ops.conv3x3_ex(h.view(n, H, W, cin), r["w1"], bias=r["bias1"], col_stats=stats_ok,
                            out_dtype=torch.bfloat16 if stats_ok else torch.float32)
        h2 = ops.groupnorm(c1.out, None, n, r["g2"], r["b2"], self.eps, True, stats1=c1.col_stats)
        o32 = want != "bf16"
        kw = dict(bias=r["bias2"], col_stats=o32 and stats_ok, out_dtype=torch.float32 if o32 else torch.bfloat16)
        if r["sc"]:
            c2 = ops.conv3x3_ex(h2.view(n, H, W, cout), r["w2"], a2=raw, **kw)
        else:
            c2 = ops.conv3x3_ex(h2.view(n, H, W, cout), r["w2"], residual=t1, **kw)
        return (None, None, c2.out) if want == "bf16" else (c2.out, c2.col_stats, None)

    def _attention(self, a: dict, x, n: int, L: int):
        """diffusers 0.10.2 AttentionBlock (one head of width C): x + proj(softmax(Q K^T / sqrt C) V)."""
        xt, xs = x[:2]
        C = xt.shape[1]
        bf = torch.bfloat16
        hn = ops.groupnorm(xt, None, n, a["g"], a["b"], self.eps, False, stats1=xs)              # [n*L, C] bf16
        q = ops.gemm(hn, a["wq"], bias=a["bq"], out_dtype=bf)
        k = ops.gemm(hn, a["wk"], bias=a["bk"], out_dtype=bf)
        scores = torch.empty((n * L, L), device=xt.device, dtype=torch.float32)
        for i in range(n):
            ops.gemm(q[i * L:(i + 1) * L], k[i * L:(i + 1) * L], out=scores[i * L:(i + 1) * L])
        p = ops.softmax_rows(scores, 1.0 / math.sqrt(C))
        o = torch.empty((n * L, C), device=xt.device, dtype=bf)
        for i in range(n):
            vt = ops.gemm(a["wv"], hn[i * L:(i + 1) * L], out_dtype=bf)                           # V^T [C, L] without a transpose pass
            ops.gemm(p[i * L:(i + 1) * L], vt, out=o[i * L:(i + 1) * L])
        r = ops.gemm_ex(o, a["wp"], bias=a["bp"], residual=xt, col_stats=(L % 32 == 0))
        return (r.out, r.col_stats, None)

    def _mid(self, m: dict, x, n: int, H: int, W: int):
        x = self._resnet(m["res"][0], x, n, H, W)
        x = self._attention(m, x, n, H * W)
        return self._resnet(m["res"][1], x, n, H, W)

    def _upsample_conv(self, up: tuple, x16: torch.Tensor, n: int, h: int, w: int):
        """Upsample2D: nearest 2x + conv3x3 as four 2x2-tap phase convs on the low-res image (packing.pack_upsample_phases)."""
        wu, bu, phases = up
        C = x16.shape[1]
        M = n * 4 * h * w
        img = x16.view(n, h, w, C)
        if (w & (w - 1)) == 0 and (h * w) % 32 == 0:
            out = torch.empty((M, wu.shape[0]), device=x16.device, dtype=torch.float32)
            st = torch.empty((M // 32, wu.shape[0], 2), device=x16.device, dtype=torch.float32)
            done = True
            for ph in range(4):
                if ops.gemm_ex(None, phases[ph], x_img=img, conv_taps=(2, 2, (ph & 1) - 1, (ph >> 1) - 1), up_phase=1 + ph, bias=bu,
                               out=out, col_stats=st) is None:
                    done = False
                    break
            if done:
                return (out, st, None)
        uc = ops.conv3x3_ex(ops.upsample2x(x16.float(), n, h, w), wu, bias=bu, col_stats=(4 * h * w) % 32 == 0)
        return (uc.out, uc.col_stats, None)

    # ------------------------------------------------------------------ decode
    @torch.no_grad()
    def _decode_chunk(self, pk: dict, z: torch.Tensor) -> torch.Tensor:
        n, _, h, w = z.shape
        rows = z.permute(0, 2, 3, 1).reshape(-1, 4).contiguous()                           # post_quant_conv: a 4x4 matrix per pixel
        z2 = ops.small_linear(rows, pk["pq_w"], pk["pq_b"]).reshape(n, h, w, 4).permute(0, 3, 1, 2).contiguous()
        x = ops.conv_in(z2.view(n, 4, 1, h, w), pk["dec_in_w"], pk["dec_in_b"], col_stats=True) + (None,)
        x = self._mid(pk["dec_mid"], x, n, h, w)
        for blk in pk["dec_up"]:
            nres = len(blk["res"])
            for j, r in enumerate(blk["res"]):
                x = self._resnet(r, x, n, h, w, want="bf16" if (blk["up"] is not None and j == nres - 1) else "f32")
            if blk["up"] is not None:
                x = self._upsample_conv(blk["up"], x[2], n, h, w)
                h, w = 2 * h, 2 * w
        y = ops.groupnorm(x[0], None, n, pk["dec_no_g"], pk["dec_no_b"], self.eps, True, out_dtype=torch.float32, stats1=x[1])
        return ops.conv_out(y, pk["dec_out_w"], pk["dec_out_b"], n, 1, h, w).reshape(n, -1, h, w)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """`vae.decode(z).sample` (utils/ddim_sampling_utils.py:39): z (n, 4, h, w) latents -> (n, 3, 8h, 8w) images."""
        if z.dim() != 4 or z.shape[1] != self.cfg["latent_channels"]:
            raise ValueError(f"expected latents (n, {self.cfg['latent_channels']}, h, w), got {tuple(z.shape)}")
        pk = self._packed or self._pack()
        with torch.cuda.device(self.device):
            z = z.to(device=self.device, dtype=torch.float32).contiguous()
            outs = [self._decode_chunk(pk, z[i: i + self.images_per_chunk]) for i in range(0, z.shape[0], self.images_per_chunk)]
        sample = torch.cat(outs) if len(outs) > 1 else outs[0]
        return SimpleNamespace(sample=sample) if return_dict else (sample,)

    # ------------------------------------------------------------------ encode
    @torch.no_grad()
    def _encode_chunk(self, pk: dict, img: torch.Tensor) -> torch.Tensor:
        n, cin, H, W = img.shape
        x4 = torch.zeros((n, 4, 1, H, W), device=img.device, dtype=torch.float32)
        x4[:, :cin, 0] = img
        x = ops.conv_in(x4, pk["enc_in_w"], pk["enc_in_b"], col_stats=True) + (None,)
        h, w = H, W
        for blk in pk["enc_down"]:
            nres = len(blk["res"])
            for j, r in enumerate(blk["res"]):
                # the stride-2 conv reads bf16 and nothing else reads the block's last tensor
                x = self._resnet(r, x, n, h, w, want="bf16" if (blk["down"] is not None and j == nres - 1) else "f32")
            if blk["down"] is not None:
                wd, bd = blk["down"]
                # Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) + stride-2 conv without padding = tap offsets (0, 0), OOB zero fill
                dn = ops.gemm_ex(None, wd, x_img=x[2].view(n, h, w, x[2].shape[1]), conv_stride=2, conv_taps=(3, 3, 0, 0), bias=bd,
                                 col_stats=True)
                if dn is None:
                    raise ValueError(f"AutoencoderKL.encode: image size {H}x{W} is outside the conv kernel's tiling (multiples of 64 work)")
                x = (dn.out, dn.col_stats, None)
                h, w = h // 2, w // 2
        x = self._mid(pk["enc_mid"], x, n, h, w)
        y = ops.groupnorm(x[0], None, n, pk["enc_no_g"], pk["enc_no_b"], self.eps, True, out_dtype=torch.float32, stats1=x[1])
        parts = [ops.conv_out(y, wp, bp, n, 1, h, w).reshape(n, -1, h, w) for wp, bp in zip(pk["enc_out_w"], pk["enc_out_b"])]
        m = torch.cat(parts, 1)                                                         # (n, 8, h, w)
        rows = m.permute(0, 2, 3, 1).reshape(-1, m.shape[1]).contiguous()               # quant_conv: an 8x8 matrix per pixel
        return ops.small_linear(rows, pk["q_w"], pk["q_b"]).reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """`vae.encode(x).latent_dist` (inference.py:186): x (n, 3, H, W) images in [-1, 1] -> DiagonalGaussianDistribution."""
        if x.dim() != 4 or x.shape[1] != self.cfg["in_channels"] or x.shape[2] % 8 or x.shape[3] % 8:
            raise ValueError(f"expected images (n, {self.cfg['in_channels']}, H, W) with H, W multiples of 8, got {tuple(x.shape)}")
        pk = self._packed or self._pack()
        with torch.cuda.device(self.device):
            x = x.to(device=self.device, dtype=torch.float32).contiguous()
            step = max(1, self.images_per_chunk // 2)
            outs = [self._encode_chunk(pk, x[i: i + step]) for i in range(0, x.shape[0], step)]
        dist = DiagonalGaussianDistribution(torch.cat(outs) if len(outs) > 1 else outs[0])
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, generator: Optional[torch.Generator] = None):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator) if sample_posterior else posterior.mode()
        return self.decode(z)
