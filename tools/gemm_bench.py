#!/usr/bin/env python
"""Per-shape micro-benchmark of the tcgen05 GEMM / conv kernel at the shapes of one UNet evaluation of the bench
workload (UNet batch 16, 16 frames, 32x32 latents).  CUDA-event timing, L2 flushed between repetitions.

    python tools/gemm_bench.py [--reps 5] [--sweep]      (--sweep: also try every legal tile width per shape)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seervideoldm_b200 import ops  # noqa: E402

DEV = "cuda"
B, F = 16, 16


def shapes():
    out = []
    for lvl, (C, h) in enumerate([(320, 32), (640, 16), (1280, 8), (1280, 4)]):
        M = B * F * h * h
        out.append((f"L{lvl} linear CxC +res f32", dict(kind="lin", M=M, N=C, K=C, res=True, out="f32")))
        out.append((f"L{lvl} proj_in (no res) f32", dict(kind="lin", M=M, N=C, K=C, res=False, out="f32")))
        out.append((f"L{lvl} qkv bf16", dict(kind="lin", M=M, N=3 * C, K=C, res=False, out="bf16")))
        out.append((f"L{lvl} q bf16", dict(kind="lin", M=M, N=C, K=C, res=False, out="bf16")))
        out.append((f"L{lvl} geglu", dict(kind="geglu", M=M, N=8 * C, K=C)))
        out.append((f"L{lvl} ff_out +res", dict(kind="lin", M=M, N=C, K=4 * C, res=True, out="f32")))
        out.append((f"L{lvl} conv3x3 C->C +res", dict(kind="conv", n_img=B * F, H=h, Cin=C, Cout=C, res=True)))
    out.append(("L0 conv3x3 960->320 + sc", dict(kind="conv", n_img=B * F, H=32, Cin=960, Cout=320, res=False, K2=960)))
    out.append(("L1 conv3x3 1920->640 + sc", dict(kind="conv", n_img=B * F, H=16, Cin=1920, Cout=640, res=False, K2=1920)))
    out.append(("L2 conv3x3 2560->1280 + sc", dict(kind="conv", n_img=B * F, H=8, Cin=2560, Cout=1280, res=False, K2=2560)))
    out.append(("L0 upsample conv 640->640", dict(kind="conv", n_img=B * F, H=32, Cin=640, Cout=640, res=False)))
    return out


def run_one(cfg, reps, flush):
    g = torch.Generator(device=DEV).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    if cfg["kind"] in ("lin", "geglu"):
        M, N, K = cfg["M"], cfg["N"], cfg["K"]
        a = rnd(M, K).bfloat16()
        w = (rnd(N, K) * K ** -0.5).bfloat16()
        bias = rnd(N)
        flops = 2.0 * M * N * K
        if cfg["kind"] == "geglu":
            fn = lambda: ops.gemm_ex(a, w, bias=bias, geglu=True)
            nbytes = M * K * 2 + N * K * 2 + M * (N // 2) * 2
        else:
            res = rnd(M, N) if cfg["res"] else None
            od = torch.float32 if cfg["out"] == "f32" else torch.bfloat16
            o = torch.empty(M, N, device=DEV, dtype=od)
            fn = lambda: ops.gemm_ex(a, w, bias=bias, residual=res, out=o)
            nbytes = M * K * 2 + N * K * 2 + M * N * o.element_size() + (M * N * 4 if res is not None else 0)
    else:
        n_img, H, Cin, Cout = cfg["n_img"], cfg["H"], cfg["Cin"], cfg["Cout"]
        K2 = cfg.get("K2", 0)
        M = n_img * H * H
        x = rnd(n_img, H, H, Cin).bfloat16()
        w = (rnd(Cout, 9 * Cin + K2) * (9 * Cin) ** -0.5).bfloat16()
        a2 = rnd(M, K2).bfloat16() if K2 else None
        bias = rnd(Cout)
        res = rnd(M, Cout) if cfg["res"] else None
        o = torch.empty(M, Cout, device=DEV)
        fn = lambda: ops.conv3x3_ex(x, w, a2=a2, bias=bias, residual=res, out=o)
        flops = 2.0 * M * Cout * (9 * Cin + K2)
        nbytes = M * (Cin + K2) * 2 + w.numel() * 2 + M * Cout * 4 * (2 if res is not None else 1)
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    return dict(ms=ms, tflops=flops / ms / 1e9, gbs=nbytes / ms / 1e6, gflop=flops / 1e9, mbytes=nbytes / 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="comma-separated substrings: run only the shapes whose name contains one")
    ap.add_argument("--cg", type=int, default=0, help="force CTA-group size 1 or 2 (tuning hook SEER_GEMM_CG)")
    args = ap.parse_args()
    if args.cg:
        ops.set_tuning("SEER_GEMM_CG", args.cg)
    flush = torch.empty(256 * 1024 * 1024, device=DEV, dtype=torch.uint8)
    rows = []
    for name, cfg in shapes():
        if args.only and not any(k in name for k in args.only.split(",")):
            continue
        variants = [0]
        if args.sweep:
            N = cfg.get("N", cfg.get("Cout"))
            cands = [256, 128] if cfg["kind"] == "geglu" else [320, 256, 192, 160, 128]
            variants = [0] + [bn for bn in cands if N % bn == 0]
        for bn in variants:
            ops.set_tuning("SEER_GEMM_BN", bn if bn else None)
            try:
                r = run_one(cfg, args.reps, flush)
            except Exception as e:  # noqa: BLE001
                print(f"{name:34s} BN={bn or 'auto':>4}  FAILED: {e}")
                continue
            r.update(name=name, bn=bn or "auto")
            rows.append(r)
            print(f"{name:34s} BN={r['bn']!s:>4}  {r['ms'] * 1e3:9.1f} us  {r['tflops']:7.1f} TF/s  {r['gbs']:7.0f} GB/s   "
                  f"({r['gflop']:.0f} GFLOP, {r['mbytes']:.0f} MB)", flush=True)
    ops.set_tuning("SEER_GEMM_BN", None)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
