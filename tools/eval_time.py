#!/usr/bin/env python
"""ms per CUDA-graph replay of ONE UNet evaluation at the bench shape (UNet batch 16 = 8 clips under CFG, 16 frames, 32x32
latents) — the A/B timer for kernel changes (31 replays = one bench step).

    python tools/eval_time.py [--clips 8] [--frames 16] [--iters 31] [--eager]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import SeerUNet  # noqa: E402
from seervideoldm_b200.graph import GraphedUNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--latent", type=int, default=32)
ap.add_argument("--iters", type=int, default=31)
ap.add_argument("--eager", action="store_true")
args = ap.parse_args()

net = SeerUNet(sample_size=32, cross_attention_dim=768)
with torch.no_grad():
    for n, p in net.named_parameters():
        if n.endswith("proj_out.weight"):
            p.normal_(std=0.02)
net = net.cuda().eval()
B = 2 * args.clips
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, args.frames, args.latent, args.latent, generator=g).cuda()
c = torch.randn(B, args.frames, 77, 768, generator=g).cuda()
t = torch.full((B,), 496, device="cuda")
fn = (lambda: net(x, t, c)) if args.eager else (lambda gr=GraphedUNet(net, x, t, c, 0): gr(x, t, c))
for _ in range(8):
    fn()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    best = min(best, ms)
    print(f"rep {rep}: {ms:.3f} ms / evaluation ({'eager' if args.eager else 'graph'})")
print(f"best {best:.3f} ms / evaluation  -> {args.clips / (31 * best * 1e-3):.3f} clips/s")
