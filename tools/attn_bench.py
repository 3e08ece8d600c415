#!/usr/bin/env python
"""Attention micro-benchmark at the bench-shape launches (UNet batch 16, 16 frames): CUDA-event time per launch,
TF/s and exp2 rate against the MUFU roofline (16 ex2 / clk / SM).

    python tools/attn_bench.py [--iters 20]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--latent", type=int, default=32, help="level-0 latent side (64 = BASELINE.json config 5, the 512x512 attention stress)")
args = ap.parse_args()
B, F, heads = args.batch, args.frames, 8
dev = "cuda"
cases = []
for d, h in ((40, args.latent), (80, args.latent // 2), (160, args.latent // 4)):
    C = heads * d
    M = B * F * h * h
    qkv = torch.randn(M, 3 * C, device=dev).bfloat16()
    kv = torch.randn(B * F * 77, 2 * C, device=dev).bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    cases.append((f"spatial d={d} L={h*h}", lambda q=q, k=k, v=v, h=h: ops.attention(q, k, v, mode=ops.ATTN_SPATIAL, heads=heads, n_outer=B * F, Lq=h * h, Lk=h * h),
                  4.0 * (h * h) ** 2 * d * B * F * heads, float((h * h) ** 2) * B * F * heads))
    cases.append((f"cross   d={d} Lq={h*h} Lk=77", lambda q=q, kv=kv, C=C, h=h: ops.attention(q, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=B * F, Lq=h * h, Lk=77),
                  4.0 * h * h * 77 * d * B * F * heads, float(h * h * 77) * B * F * heads))
    ws = 0 if h <= 4 else (8 if h // 8 >= 4 else 4)
    L = F * (ws * ws if ws else h * h)
    nprob = B * heads * ((h // ws) ** 2 if ws else 1)
    cases.append((f"scta    d={d} L={L}", lambda q=q, k=k, v=v, h=h: ops.attention(q, k, v, mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=F, H=h, W=h),
                  4.0 * L * (L + 1) / 2 * d * nprob, L * (L + 1) / 2 * nprob))
for name, fn, flops, exps in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"{name:32s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TF/s  {exps / ms / 1e9:7.2f} Texp/s (algorithmic; MUFU peak ~3.8 at 1.6 GHz)")
