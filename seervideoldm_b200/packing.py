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
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] bf16 with K order [Cin/64][ky][kx][64] (one K-block per (64-channel slab, tap),
    the 9 taps of a slab adjacent so the implicit-GEMM conv re-reads a slab while it is hot in L2); an optional 1x1
    shortcut weight [Cout, Csc, 1, 1] is appended along K (fused ResNet shortcut)."""
    cout, cin = w.shape[:2]
    if cin % 64:
        raise ValueError("pack_conv3x3: Cin must be a multiple of 64")
    p = w.permute(0, 2, 3, 1).reshape(cout, 9, cin // 64, 64).permute(0, 2, 1, 3).reshape(cout, 9 * cin)
    if shortcut is not None:
        p = torch.cat([p, shortcut.reshape(cout, -1)], dim=1)
    return p.to(torch.bfloat16).contiguous()


def pack_upsample_phases(w: torch.Tensor) -> list:
    """Upsample3D = nearest 2x then conv3x3 (resnet.py:47-61).  Output pixel (2y+py, 2x+px) only ever sees the 2x2 low-res
    neighbourhood {y+py-1, y+py} x {x+px-1, x+px}, each low-res pixel through the SUM of the 3x3 taps that land on its
    duplicates: rows (py=0) [w0 | w1+w2], (py=1) [w0+w1 | w2], same along x.  Returns the four phase weights (index 2 py + px),
    each [Cout, 4*Cin] bf16 in K order [Cin/64][ty][tx][64] — 4/9 of the FLOPs of convolving the upsampled image."""
    cout, cin = w.shape[:2]
    if cin % 64:
        raise ValueError("pack_upsample_phases: Cin must be a multiple of 64")
    w = w.float()
    rows = {0: (w[:, :, 0], w[:, :, 1] + w[:, :, 2]), 1: (w[:, :, 0] + w[:, :, 1], w[:, :, 2])}      # [Cout, Cin, kx] per ty
    out = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for ty in (0, 1):
                r = rows[py][ty]
                cols = (r[:, :, 0], r[:, :, 1] + r[:, :, 2]) if px == 0 else (r[:, :, 0] + r[:, :, 1], r[:, :, 2])
                taps += [cols[0], cols[1]]
            p = torch.stack(taps, dim=1)                                   # [Cout, 4, Cin]
            p = p.reshape(cout, 4, cin // 64, 64).permute(0, 2, 1, 3).reshape(cout, 4 * cin)
            out.append(p.to(torch.bfloat16).contiguous())
    return out


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


def fold_layernorm(w: torch.Tensor, b: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                   geglu: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fold LayerNorm(gamma, beta) into the Linear(w [N, C], b) that consumes it:

        Linear(LN(x)) = rstd * (x @ (w*gamma).T - mean * colsum) + (w @ beta + b),   colsum[n] = sum_k (w*gamma)[n, k]

    so the GEMM runs on the RAW activations and applies the per-row (mean, rstd) in its epilogue (SeerGemmDesc LN fold;
    reference: attention.py:237,244,311,322-323 apply nn.LayerNorm in front of attn1 / attn2 / ff).  Returns
    (w*gamma as bf16, colsum fp32 of the ROUNDED weights, folded bias fp32); rows in GEGLU block order if `geglu`."""
    w = w.float()
    wf = w * gamma.float()[None, :]
    bias = w @ beta.float()
    if b is not None:
        bias = bias + b.float()
    if geglu:
        perm = geglu_permutation(w.shape[0] // 2, device=w.device)
        wf, bias = wf[perm], bias[perm]
    w16 = wf.to(torch.bfloat16).contiguous()
    return w16, w16.float().sum(1).contiguous(), bias.contiguous()


# ---- fp32-parity path: error-compensated bf16 operand pairs (csrc/fp32_path.cu) ----
def split3_weight(w: torch.Tensor) -> torch.Tensor:
    """[N, K, ...] fp32 -> [N, 3K, ...] fp32 holding [w_hi | w_lo | w_hi] along dim 1 (both halves exactly representable in
    bf16), the partner of the activation split [a_hi | a_hi | a_lo]: the 3K-long dot product is
    a_hi.w_hi + a_hi.w_lo + a_lo.w_hi.  Feed the result to the ordinary pack_* functions."""
    w = w.float()
    hi = w.to(torch.bfloat16).float()
    lo = (w - hi).to(torch.bfloat16).float()
    return torch.cat([hi, lo, hi], dim=1)
