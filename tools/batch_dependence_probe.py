#!/usr/bin/env python
"""Per-op bisect of batch dependence: one eager UNet evaluation of clip 0 alone (UNet batch 1) and inside a batch of 2,
comparing the first clip's slice of EVERY op output bit for bit, plus a run-to-run determinism check.

    python tools/batch_dependence_probe.py [--frames 16] [--latent 32]
"""
import argparse
import dataclasses
import inspect
import os
import sys
from collections import Counter

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import SeerUNet, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--latent", type=int, default=32)
args = ap.parse_args()

net = SeerUNet(sample_size=32, cross_attention_dim=768)
with torch.no_grad():
    for n, p in net.named_parameters():
        if n.endswith("proj_out.weight"):
            p.normal_(std=0.02)
net = net.cuda().eval()
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 4, args.frames, args.latent, args.latent, generator=g).cuda()
c = torch.randn(2, args.frames, 77, 768, generator=g).cuda()
t = torch.full((2,), 496, device="cuda")


def tensors(res):
    if torch.is_tensor(res):
        yield res
    elif dataclasses.is_dataclass(res):
        for f in dataclasses.fields(res):
            yield from tensors(getattr(res, f.name))
    elif isinstance(res, (tuple, list)):
        for r in res:
            yield from tensors(r)
    elif isinstance(res, ops.GemmOut):
        yield from tensors((res.out, res.out16, res.col_stats))
        if res.row_stats is not None:                     # [parts, M, 2]: make it row-major like everything else
            yield res.row_stats.permute(1, 0, 2)


LOG, B = [], 1


def wrap(name, fn):
    def inner(*a, **k):
        res = fn(*a, **k)
        for i, tt in enumerate(tensors(res)):
            if tt.dim() >= 1 and tt.shape[0] % B == 0 and tt.shape[0] >= B:
                shape_tag = tuple(tt.shape[1:])
                LOG.append((f"{name}#{i} {tuple(tt.shape[:1])[0] // B}x{shape_tag} {str(tt.dtype)[6:]}", tt[: tt.shape[0] // B].detach().clone()))
        return res
    return inner


for nm, fn in list(vars(ops).items()):
    if not nm.startswith("_") and inspect.isfunction(fn) and fn.__module__ == ops.__name__:
        setattr(ops, nm, wrap(nm, fn))

def logged(xx, tt, cc):
    global LOG
    net._kv_key = None               # every logged run recomputes the text K/V, so the op sequences line up
    LOG = []
    y = net(xx, tt, cc).clone()
    return y, LOG


with torch.no_grad():
    net(x, t, c)                     # packs weights, sets function attributes
    B = 2
    y2, log2 = logged(x, t, c)
    y2b, log2b = logged(x, t, c)
    B = 1
    y1, log1 = logged(x[:1].contiguous(), t[:1], c[:1].contiguous())
torch.cuda.synchronize()

print(f"shape: UNet batch 2 vs 1, {args.frames} frames, {args.latent}x{args.latent} latents; {len(log2)} recorded op outputs per evaluation")
print(f"run-to-run determinism at batch 2: output bit-identical {torch.equal(y2, y2b)}; "
      f"op outputs differing between the two runs: {sum(not torch.equal(a[1], b[1]) for a, b in zip(log2, log2b))}")
print(f"clip 0 alone vs inside the batch: output bit-identical {torch.equal(y1[0], y2[0])}, "
      f"rel-L2 {float((y1[0] - y2[0]).norm() / y2[0].norm()):.3e}")
if len(log1) != len(log2):
    print(f"op sequences differ in length: {len(log1)} vs {len(log2)}")
bad = Counter()
first = None
for i, ((n1, a), (n2, b)) in enumerate(zip(log1, log2)):
    if a.shape != b.shape:
        bad[n1 + " (shape)"] += 1
        continue
    if not torch.equal(a, b):
        if first is None:
            d = (a.float() - b.float())
            first = (i, n1, float(d.abs().max()), float(d.norm() / (b.float().norm() + 1e-30)), int((a != b).sum()), a.numel())
        bad[n1.split("#")[0]] += 1
if first is None:
    print("every op output of clip 0 is bit-identical")
else:
    i, n, mx, rel, cnt, tot = first
    print(f"first differing op output: #{i} {n}: {cnt} of {tot} elements differ, max abs {mx:.3e}, rel-L2 {rel:.3e}")
    prev = [nm for nm, _ in log1[max(0, i - 6): i]]
    print("  preceding op outputs (all identical): " + " | ".join(prev))
    print("differing outputs per op (downstream of the first one, most are consequences):")
    for k, v in bad.most_common(12):
        print(f"  {k:30s} {v}")
