"""One eager UNet evaluation at the bench shape inside cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--clips 8] [--frames 16]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import SeerUNet, ops  # noqa: E402
from seervideoldm_b200.config import sd15_config  # noqa: E402
from seervideoldm_b200.weights import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--latent", type=int, default=32)
ap.add_argument("--fast-init", action="store_true", help="skip the seeded weight factory (profiling only)")
ap.add_argument("--list-gemm", default=None, help="write the ordered GEMM / conv launch list of the profiled evaluation (name, FLOPs, algorithmic bytes) as JSON")
args = ap.parse_args()

net = SeerUNet(sample_size=32, cross_attention_dim=768)
if not args.fast_init:
    net.load_state_dict(random_state_dict(sd15_config(sample_size=32), seed=0), strict=True)
else:
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
for _ in range(2):
    net(x, t, c)
torch.cuda.synchronize()
if args.list_gemm:
    import json
    ops.PROFILE = []
    net(x, t, c)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    rows = [dict(name=p[0], flops=p[1], bytes=p[4], ms=p[2].elapsed_time(p[3])) for p in prof if p[0].startswith(("gemm ", "conv"))]
    with open(args.list_gemm, "w") as f:
        json.dump(rows, f, indent=0)
    print(f"wrote {len(rows)} GEMM / conv launches to {args.list_gemm}")
torch.cuda.profiler.start()
net(x, t, c)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
