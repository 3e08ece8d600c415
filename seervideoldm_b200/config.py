"""Architecture constants of the Seer UNet (SD-1.5 inflated).

Mirrors the constructor arguments of the reference's `SeerUNet`
(/root/reference/seer/models/unet_3d_condition.py:64-84); the reference hard-overrides the block
types and `downsample_padding` (ibid. :90-92), so they are not configurable here either.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


@dataclass(frozen=True)
class UNetConfig:
    sample_size: int | None = None
    in_channels: int = 4
    out_channels: int = 4
    center_input_sample: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D")
    up_block_types: Tuple[str, ...] = ("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D")
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    downsample_padding: int = 1
    mid_block_scale_factor: float = 1
    act_fn: str = "silu"
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    cross_attention_dim: int = 1280
    attention_head_dim: int = 8      # used as the number of heads (unet_3d_blocks.py:170-177)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def heads(self) -> int:
        return self.attention_head_dim


def sd15_config(**over) -> UNetConfig:
    """The configuration the reference loads from runwayml/stable-diffusion-v1-5 `unet/config.json`
    (inference.py:82-87): cross_attention_dim 768, sample_size 64."""
    kw = dict(sample_size=64, cross_attention_dim=768)
    kw.update(over)
    return UNetConfig(**kw)
