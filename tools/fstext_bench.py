#!/usr/bin/env python
"""FSTextTransformer timing at the shipped configuration (num_frames 16, num_layers 8; inference.py:88):
ms per forward for a batch of clips, both precisions, CUDA events.  The module runs once per clip, upstream of the loop.

    python tools/fstext_bench.py [--clips 8] [--iters 10]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import FSTextTransformer  # noqa: E402
from seervideoldm_b200.weights import random_fstext_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
m = FSTextTransformer(num_frames=16, num_layers=8)
m.load_state_dict(random_fstext_state_dict(16, 8, seed=0), strict=True)
m = m.cuda().eval()
ctx = torch.randn(args.clips, 77, 768, device="cuda")
# per clip: 8 layers x (qkv/o/q2/kv2/o2 + 2 GEGLU FFs) on F*77 = 1232 tokens of width 768
flop = args.clips * 8 * 2 * (16 * 77) * 768 * (3 * 768 + 768 + 768 + 768 + 8 * 768 + 4 * 768 + 3 * 768 + 768 + 8 * 768 + 4 * 768)
for prec in ("bf16", "fp32"):
    m.set_precision(prec)
    for _ in range(3):
        y = m(ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        y = m(ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"FSText {prec}: {ms:.3f} ms per forward of {args.clips} clips (16 frames x 77 tokens), "
          f"{flop / ms / 1e9:.1f} TF/s on the linear layers, output {tuple(y.shape)}")
