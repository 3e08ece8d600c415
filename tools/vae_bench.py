#!/usr/bin/env python
"""AutoencoderKL decode / encode on the kernel library (seervideoldm_b200/vae.py): time per clip and per-op breakdown.

The reference decodes every sampled clip frame by frame through diffusers' fp16 VAE (utils/ddim_sampling_utils.py:36-41:
`vae.decode(latents / 0.18215)`, 16 images of 32x32x4 -> 256x256x3 per clip) and encodes the conditioning frames
(:20-27).  This prints what that hand-off costs next to the 2.26 s sampling pass.

    python tools/vae_bench.py [--clips 8] [--frames 16] [--latent 32] [--iters 5]
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seervideoldm_b200 import AutoencoderKL, ops  # noqa: E402
from seervideoldm_b200.vae import random_vae_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--latent", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()

vae = AutoencoderKL()
vae.load_state_dict(random_vae_state_dict(seed=0), strict=True)
vae = vae.cuda().eval()
n = args.clips * args.frames
g = torch.Generator().manual_seed(0)
z = torch.randn(n, 4, args.latent, args.latent, generator=g).cuda()
img = torch.randn(n, 3, 8 * args.latent, 8 * args.latent, generator=g).clamp_(-1, 1).cuda()


def timed(fn, arg):
    for _ in range(2):
        fn(arg)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(arg)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ops.PROFILE = []
    fn(arg)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    return best, prof


for name, fn, arg in (("decode", lambda a: vae.decode(a).sample, z), ("encode", lambda a: vae.encode(a).latent_dist.mode(), img)):
    ms, prof = timed(fn, arg)
    flops = sum(p[1] for p in prof)
    in_ops = sum(p[2].elapsed_time(p[3]) for p in prof)
    print(f"VAE {name}: {n} images ({args.clips} clips x {args.frames} frames, {args.latent}x{args.latent} latents <-> "
          f"{8 * args.latent}x{8 * args.latent} px): {ms:.1f} ms = {ms / args.clips:.2f} ms per clip, {flops / 1e12:.2f} TFLOP in GEMM/conv/"
          f"attention launches -> {flops / ms / 1e9:.0f} TF/s whole pass ({len(prof)} launches, {in_ops:.1f} ms inside op events)")
    agg = collections.OrderedDict()
    for nm, fl, a, b, *_ in prof:
        d = agg.setdefault(nm.split(" M=")[0] if nm.startswith("conv") or nm.startswith("gemm") else nm, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += fl
    for nm, (cnt, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        tf = fl / t / 1e9 if t > 0 and fl > 0 else 0.0
        print(f"    {nm:40s} n={cnt:4d} {t:8.2f} ms {100 * t / in_ops:5.1f}%  {tf:7.1f} TF/s")
