#!/usr/bin/env python
"""The HBM-/epilogue-bound launches of the 32x32 level (UNet batch 16, 16 frames: M = 262144 tokens, C = 320) in
isolation, with the same epilogue options the UNet uses, inside cudaProfilerStart/Stop for ncu:

  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/l0 python tools/probe_l0.py
  python tools/probe_l0.py --time        (CUDA-event timing, L2 flushed between launches)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--time", action="store_true")
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--only", default=None)
args = ap.parse_args()

C, h = [(320, 32), (640, 16), (1280, 8), (1280, 4)][args.level]
B, F, heads = 16, 16, 8
M = B * F * h * h
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
x16 = rnd(M, C).bfloat16()
x32 = rnd(M, C)
hid = rnd(M, 4 * C).bfloat16()
w = lambda n, k: (rnd(n, k) * k ** -0.5).bfloat16()
w_cc, w_qkv, w_ff1, w_ff2 = w(C, C), w(3 * C, C), w(8 * C, C), w(C, 4 * C)
b_c, b_qkv, b_ff1 = rnd(C), rnd(3 * C), rnd(8 * C)
cs_c, cs_qkv, cs_ff1 = rnd(C), rnd(3 * C), rnd(8 * C)
r0 = ops.gemm_ex(x16, w_cc, bias=b_c, also_bf16=True, row_stats=True)
qkv = ops.gemm_ex(x16, w_qkv, bias=b_qkv, out_dtype=torch.bfloat16, ln=(r0.row_stats, cs_qkv, 1e-5)).out
kv = rnd(B * F * 77, 2 * C).bfloat16()

cases = {
    "pin  (CxC, f32+bf16 out, row stats)": lambda: ops.gemm_ex(x16, w_cc, bias=b_c, also_bf16=True, row_stats=True),
    "o1   (CxC, +res f32, f32+bf16 out, row stats)": lambda: ops.gemm_ex(x16, w_cc, bias=b_c, residual=x32, also_bf16=True, row_stats=True),
    "pout (CxC, +res f32, f32 out, col stats)": lambda: ops.gemm_ex(x16, w_cc, bias=b_c, residual=x32, col_stats=True),
    "q2   (CxC, LN fold, bf16 out)": lambda: ops.gemm_ex(x16, w_cc, bias=b_c, out_dtype=torch.bfloat16, ln=(r0.row_stats, cs_c, 1e-5)),
    "qkv  (Cx3C, LN fold, bf16 out)": lambda: ops.gemm_ex(x16, w_qkv, bias=b_qkv, out_dtype=torch.bfloat16, ln=(r0.row_stats, cs_qkv, 1e-5)),
    "qkv  (Cx3C, bias only, bf16 out)": lambda: ops.gemm_ex(x16, w_qkv, bias=b_qkv, out_dtype=torch.bfloat16),
    "ff1  (Cx8C GEGLU, LN fold)": lambda: ops.gemm_ex(x16, w_ff1, bias=b_ff1, geglu=True, ln=(r0.row_stats, cs_ff1, 1e-5)),
    "ff1  (Cx8C GEGLU, bias only)": lambda: ops.gemm_ex(x16, w_ff1, bias=b_ff1, geglu=True),
    "ff2  (4CxC, +res f32, bf16 out)": lambda: ops.gemm_ex(hid, w_ff2, bias=b_c, residual=x32, out_dtype=torch.bfloat16),
    "attn spatial": lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SPATIAL, heads=heads, n_outer=B * F, Lq=h * h, Lk=h * h),
    "attn scta": lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=F, H=h, W=h),
    "attn cross": lambda: ops.attention(qkv[:, :C], kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=B * F, Lq=h * h, Lk=77),
}
if args.only:
    cases = {k: v for k, v in cases.items() if any(s in k for s in args.only.split(","))}
for fn in cases.values():
    fn()
torch.cuda.synchronize()
if args.time:
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for name, fn in cases.items():
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{name:50s} {sorted(ts)[2] * 1e3:9.1f} us", flush=True)
else:
    torch.cuda.profiler.start()
    for fn in cases.values():
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
