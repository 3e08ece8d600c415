#!/usr/bin/env python
"""Probe of single GEMM launches at benchmark shapes (CUDA events, L2 flushed between repetitions, plan printed):

    python tools/gemm_probe.py [--case qkv|qkv_rope|o1|pin|pout|ff1|all] [--reps 5] [--once]

`--once` runs each case exactly twice (warm-up + one launch) — for `ncu -k regex:gemm_tc -s 1 -c 1` captures.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seervideoldm_b200 import ops  # noqa: E402

DEV = "cuda"


def build(case, M, C):
    g = torch.Generator(device=DEV).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    a = rnd(M, C).bfloat16()
    bf = torch.bfloat16
    if case in ("qkv", "qkv_rope"):
        prod = ops.gemm_ex(a, (torch.eye(C, device=DEV) * 1.0).bfloat16(), bias=rnd(C), out_dtype=bf, row_stats=True)
        w = (rnd(3 * C, C) * C ** -0.5).bfloat16()
        colsum, bias = w.float().sum(1).contiguous(), rnd(3 * C)
        rope = None
        if case == "qkv_rope":
            T = 16384 if M % 16384 == 0 else M
            freqs = (1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32))).to(DEV)
            rope = (ops.rope_table(freqs, T), 2 * C, C // 8)
        out = torch.empty(M, 3 * C, device=DEV, dtype=bf)
        return (lambda: ops.gemm_ex(prod.out, w, bias=bias, out=out, ln=(prod.row_stats, colsum, 1e-5), rope=rope)), 2.0 * M * 3 * C * C
    w = (rnd(C, C) * C ** -0.5).bfloat16()
    bias = rnd(C)
    if case == "pin":
        out = torch.empty(M, C, device=DEV, dtype=bf)
        return (lambda: ops.gemm_ex(a, w, bias=bias, out=out, row_stats=True)), 2.0 * M * C * C
    if case == "o1":
        res = rnd(M, C).bfloat16()
        out = torch.empty(M, C, device=DEV, dtype=bf)
        return (lambda: ops.gemm_ex(a, w, bias=bias, residual=res, out=out, row_stats=True)), 2.0 * M * C * C
    if case == "pout":
        res = rnd(M, C)
        out = torch.empty(M, C, device=DEV)
        return (lambda: ops.gemm_ex(a, w, bias=bias, residual=res, out=out, col_stats=True)), 2.0 * M * C * C
    if case == "ff1":
        prod = ops.gemm_ex(a, (torch.eye(C, device=DEV) * 1.0).bfloat16(), bias=rnd(C), out_dtype=bf, row_stats=True)
        w8 = (rnd(8 * C, C) * C ** -0.5).bfloat16()
        colsum, b8 = w8.float().sum(1).contiguous(), rnd(8 * C)
        out = torch.empty(M, 4 * C, device=DEV, dtype=bf)
        return (lambda: ops.gemm_ex(prod.out, w8, bias=b8, out=out, geglu=True, ln=(prod.row_stats, colsum, 1e-5))), 2.0 * M * 8 * C * C
    raise SystemExit(f"unknown case {case}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--M", type=int, default=262144)
    ap.add_argument("--C", type=int, default=320)
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()
    cases = ["pin", "o1", "pout", "qkv", "qkv_rope", "ff1"] if args.case == "all" else args.case.split(",")
    flush = torch.empty(256 * 1024 * 1024, device=DEV, dtype=torch.uint8)
    for case in cases:
        fn, flops = build(case, args.M, args.C)
        fn()
        torch.cuda.synchronize()
        if args.once:
            flush.zero_()
            fn()
            torch.cuda.synchronize()
            print(f"{case:10s} {ops.last_gemm_kernel()}")
            continue
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{case:10s} M={args.M} C={args.C}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF/s   {ops.last_gemm_kernel()}", flush=True)


if __name__ == "__main__":
    main()
