"""Repacking of reference-layout weights into the layouts the kernels read.

All functions take CPU or CUDA fp32 tensors in the reference's shapes and return new tensors; they
run once at load time (`SeerUNet._pack`).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

GEGLU_BLOCK = 32   # value/gate column interleave granularity = the epilogue's TMEM load width


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """nn.Linear weight [N, K] is already K-major: just bf16."""
    return w.to(torch.bfloat16).contiguous()


def pack_conv1x1(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 1, 1] -> [Cout, Cin] bf16."""
    return w.reshape(w.shape[0], w.shape[1]).to(torch.bfloat16).contiguous()


def pack_conv3x3(w: torch.Tensor, shortcut: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] bf16 with K order [ky][kx][Cin] (one K-block per (tap, 64 channels));
    an optional 1x1 shortcut weight [Cout, Csc, 1, 1] is appended along K (fused ResNet shortcut)."""
    cout, cin = w.shape[:2]
    p = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    if shortcut is not None:
        p = torch.cat([p, shortcut.reshape(cout, -1)], dim=1)
    return p.to(torch.bfloat16).contiguous()


def pack_conv_out(w: torch.Tensor) -> torch.Tensor:
    """[Cout<=4, Cin, 3, 3] -> fp32 [Cout, 9, Cin]."""
    cout, cin = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(cout, 9, cin).float().contiguous()


def geglu_permutation(inner: int, device=None) -> torch.Tensor:
    """Row order for the GEGLU projection [2*inner, C]: blocks of 32 value rows followed by the matching 32 gate
    rows, so a 64-column slab of the accumulator holds (value, gate) for 32 hidden units
    (reference split: `hidden, gate = proj(x).chunk(2, -1)`, attention.py:792)."""
    assert inner % GEGLU_BLOCK == 0
    v = torch.arange(inner, device=device).reshape(-1, GEGLU_BLOCK)
    return torch.cat([v, v + inner], dim=1).reshape(-1)


def pack_geglu(w: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    perm = geglu_permutation(w.shape[0] // 2, device=w.device)
    return w[perm].to(torch.bfloat16).contiguous(), b[perm].float().contiguous()


def pack_qkv(wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
    return torch.cat([wq, wk, wv], dim=0).to(torch.bfloat16).contiguous()


def pack_kv(wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
    return torch.cat([wk, wv], dim=0).to(torch.bfloat16).contiguous()
