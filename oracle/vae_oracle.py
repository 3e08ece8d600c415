"""TEST INFRASTRUCTURE — CPU oracle of the Stable-Diffusion-1.5 `AutoencoderKL` the reference decodes latents with.

The reference calls a THIRD-PARTY network: `vae.decode(1 / 0.18215 * z).sample` (utils/ddim_sampling_utils.py:37-41) and
`vae.encode(frames).latent_dist.sample() * 0.18215` (inference.py:186-187), `vae = AutoencoderKL.from_pretrained(...,
subfolder="vae")` from **diffusers==0.10.2** (requirements.txt:8).  diffusers is not vendored under /root/reference and is not
installed in this image, so this file RESTATES the published algorithm of that version (diffusers/models/vae.py `Encoder`,
`Decoder`, `DiagonalGaussianDistribution`; unet_2d_blocks.py `DownEncoderBlock2D`, `UpDecoderBlock2D`, `UNetMidBlock2D`;
resnet.py `ResnetBlock2D`, `Downsample2D(padding=0)`, `Upsample2D`; attention.py `AttentionBlock`) in plain PyTorch with the
SD-1.5 VAE configuration (block_out_channels 128/256/512/512, 2 layers per block, 4 latent channels, 32 groups, eps 1e-6).

**Parity unpinned**: there are no golden vectors for this network in the reference and the dependency cannot be run here;
the anchors are the state-dict schema (key names / shapes of the published checkpoint, 83 653 863 parameters) and the
reference's call sites.  Only tests/ may import this module.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
GROUPS, EPS = 32, 1e-6


def _gn(sd, p, x, silu):
    y = F.group_norm(x, GROUPS, sd[p + ".weight"], sd[p + ".bias"], EPS)
    return F.silu(y) if silu else y


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """diffusers ResnetBlock2D with temb=None, output_scale_factor=1: x + conv2(silu(gn2(conv1(silu(gn1(x))))))."""
    h = _conv(sd, p + ".conv1", _gn(sd, p + ".norm1", x, True))
    h = _conv(sd, p + ".conv2", _gn(sd, p + ".norm2", h, True))
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention_block(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """diffusers 0.10.2 AttentionBlock, one head of width C: softmax(Q K^T / sqrt(C)) V, projected, + residual."""
    n, c, h, w = x.shape
    t = _gn(sd, p + ".group_norm", x, False).reshape(n, c, h * w).transpose(1, 2)
    lin = lambda name, v: F.linear(v, sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"])
    q, k, v = lin("query", t), lin("key", t), lin("value", t)
    s = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(c), dim=-1)
    o = lin("proj_attn", s @ v)
    return x + o.transpose(1, 2).reshape(n, c, h, w)


def mid_block(sd, p, x):
    x = resnet_block(sd, p + ".resnets.0", x)
    x = attention_block(sd, p + ".attentions.0", x)
    return resnet_block(sd, p + ".resnets.1", x)


def decode(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """AutoencoderKL.decode(z).sample: z (n, 4, h, w) -> (n, 3, 8h, 8w)."""
    x = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = _conv(sd, "decoder.conv_in", x)
    x = mid_block(sd, "decoder.mid_block", x)
    for i in range(4):
        for j in range(3):
            x = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    return _conv(sd, "decoder.conv_out", _gn(sd, "decoder.conv_norm_out", x, True))


def encode_moments(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """quant_conv(Encoder(x)): x (n, 3, H, W) -> moments (n, 8, H/8, W/8) = [mean | logvar]."""
    h = _conv(sd, "encoder.conv_in", x)
    for i in range(4):
        for j in range(2):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if i < 3:
            h = F.pad(h, (0, 1, 0, 1))                       # Downsample2D(padding=0): pad right / bottom, stride-2 conv without padding
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = mid_block(sd, "encoder.mid_block", h)
    h = _conv(sd, "encoder.conv_out", _gn(sd, "encoder.conv_norm_out", h, True))
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def gaussian_sample(moments: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution(moments).sample() with the given standard-normal noise."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


def decode_latents(sd: Dict[str, Tensor], samples: Tensor) -> Tensor:
    """The second half of the reference's ddim_sample (utils/ddim_sampling_utils.py:37-41): (n c f h w) latents -> pixels in [0, 1]."""
    n, c, f, h, w = samples.shape
    z = 1 / 0.18215 * samples.permute(0, 2, 1, 3, 4).reshape(n * f, c, h, w)
    x = decode(sd, z)
    x = x.reshape(n, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)
    return torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0)
