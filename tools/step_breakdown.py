#!/usr/bin/env python
"""Per-op breakdown of ONE eager UNet evaluation at the bench shape (CUDA events around every launch, in situ:
warm L2, sustained clocks after a few warm-up evaluations).  Prints time / share / TF/s per op shape.

    python tools/step_breakdown.py [--clips 8] [--frames 16] [--latent 32] [--json out.json]
"""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seervideoldm_b200 import SeerUNet, ops  # noqa: E402
from seervideoldm_b200.config import sd15_config  # noqa: E402
from seervideoldm_b200.weights import random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--warm", type=int, default=6)
    ap.add_argument("--json", default=None)
    ap.add_argument("--detail", default=None, help="print every launch whose op name contains this string (time + dispatched kernel / plan)")
    ap.add_argument("--fast-init", action="store_true", help="skip the seeded weight factory (profiling only)")
    args = ap.parse_args()
    cfg = sd15_config(sample_size=args.latent)
    net = SeerUNet(sample_size=args.latent, cross_attention_dim=768)
    if not args.fast_init:
        net.load_state_dict(random_state_dict(cfg, seed=0), strict=True)
    net = net.cuda().eval()
    if args.fast_init:                               # profiling only: live attention paths without the seeded factory
        with torch.no_grad():
            for n_, p_ in net.named_parameters():
                if n_.endswith("proj_out.weight"):
                    p_.normal_(0.0, 0.02)
    B = 2 * args.clips
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 4, args.frames, args.latent, args.latent, generator=g).cuda()
    c = torch.randn(B, args.frames, 77, 768, generator=g).cuda()
    t = torch.full((B,), 496, device="cuda")
    for _ in range(args.warm):
        net(x, t, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = []
    e0.record()
    net(x, t, c)
    e1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    wall = e0.elapsed_time(e1)
    agg = collections.OrderedDict()
    for name, flops, a, b, *_ in prof:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += flops
    tot = sum(v[1] for v in agg.values())
    print(f"eager evaluation: {wall:.2f} ms wall, {tot:.2f} ms inside op events, {len(prof)} ops")
    rows = []
    for name, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tf = fl / ms / 1e9 if ms > 0 and fl > 0 else 0.0
        rows.append(dict(op=name, n=n, ms=ms, share=ms / tot, tflops=tf))
        print(f"{name:44s} n={n:3d} {ms:8.3f} ms {100 * ms / tot:5.1f}%  {tf:7.1f} TF/s  ({ms / n * 1e3:7.1f} us each)")
    if args.detail:
        print(f"-- launches matching {args.detail!r}")
        for i, (name, flops, a, b, nbytes, *rest) in enumerate(prof):
            if args.detail in name:
                ms = a.elapsed_time(b)
                print(f"  #{i:3d} {name:40s} {ms * 1e3:8.1f} us  {flops / ms / 1e9 if ms > 0 else 0:7.1f} TF/s  {nbytes / 1e6:7.1f} MB  {rest[0] if rest else ''}")
    if args.json:
        with open(args.json, "w") as f:
            json.dump(dict(wall_ms=wall, op_ms=tot, rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
