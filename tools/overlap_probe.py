#!/usr/bin/env python
"""Does running two half-batches on two CUDA streams beat one full batch?  (HBM-bound GroupNorm / RoPE / small-K launches of one
stream could hide under tensor-bound convs of the other.)  CUDA graphs of each variant, device-timed.

    python tools/overlap_probe.py [--clips 8] [--frames 16] [--iters 5]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seervideoldm_b200 import SeerUNet  # noqa: E402
from seervideoldm_b200.graph import GraphedUNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--parts", type=int, default=2)
args = ap.parse_args()
net = SeerUNet(sample_size=32, cross_attention_dim=768)
with torch.no_grad():
    for n, p in net.named_parameters():
        if n.endswith("proj_out.weight"):
            p.normal_(std=0.02)
net = net.cuda().eval()
b = args.clips
g = torch.Generator().manual_seed(0)
x = torch.randn(b, 4, args.frames, 32, 32, generator=g).cuda()
c = torch.randn(2 * b, args.frames, 77, 768, generator=g).cuda()
t = torch.full((2 * b,), 496, device="cuda")
x_in = torch.cat([x, x])


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


full = GraphedUNet(net, x_in, t, c, 0, cfg_shared=True)
t_full = timed(lambda: full(x_in, t, c))
print(f"one graph, UNet batch {2 * b}: {t_full:.2f} ms per evaluation")

P = args.parts
per = b // P
parts = []
for i in range(P):
    sl = slice(i * per, (i + 1) * per)
    xi = torch.cat([x[sl], x[sl]])
    ci = torch.cat([c[:b][sl], c[b:][sl]]).contiguous()
    ti = t[: 2 * per]
    parts.append((GraphedUNet(net, xi, ti, ci, 0, cfg_shared=True), xi, ti, ci))
streams = [torch.cuda.Stream() for _ in range(P)]


def run_parts():
    cur = torch.cuda.current_stream()
    for s, (gph, xi, ti, ci) in zip(streams, parts):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            gph(xi, ti, ci)
    for s in streams:
        cur.wait_stream(s)


t_parts = timed(run_parts)
print(f"{P} graphs of UNet batch {2 * per} on {P} streams: {t_parts:.2f} ms per evaluation of all {b} clips  ({t_full / t_parts:.3f}x)")
t_seq = timed(lambda: [gph(xi, ti, ci) for gph, xi, ti, ci in parts])
print(f"{P} graphs of UNet batch {2 * per} back to back on one stream: {t_seq:.2f} ms")
