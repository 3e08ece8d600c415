#!/usr/bin/env python
"""CFG-branch split (parallel.py, optional latency mode) on 2 GPUs: both ranks of the pair must hold bit-identical
latents, within TOL (rel-L2) of the single-GPU `[uc; c]` run — a UNet batch of b and of 2b take different tile plans at
the full shape, so the two differ by bf16 rounding noise; the per-clip latency of both modes is printed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/cfg_branch_split_check.py [--clips 1] [--frames 16]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seervideoldm_b200 import DDIMSampler, SeerUNet  # noqa: E402
from seervideoldm_b200.config import sd15_config  # noqa: E402
from seervideoldm_b200.parallel import cfg_branch_group  # noqa: E402
from seervideoldm_b200.pipeline import ddim_sample_latents  # noqa: E402
from seervideoldm_b200.weights import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=1)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--ref-frames", type=int, default=1)
ap.add_argument("--steps", type=int, default=30, help="DDIM steps (30 = the reference's default: 31 evaluations)")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
group, branch = cfg_branch_group()

net = SeerUNet(sample_size=32, cross_attention_dim=768)
net.load_state_dict(random_state_dict(sd15_config(sample_size=32), seed=0), strict=True)
net = net.cuda().eval()

ok_all = True
TOL = 2e-2            # 31-step budget of the bf16 path is 5e-2 (BASELINE.json north_star)
for F, F1 in ((args.frames, args.ref_frames), (12, 2)):       # bench shape, then the Sthv2 shape (BASELINE.json configs[1])
    b = args.clips
    g = torch.Generator().manual_seed(1000)                      # same clips on both ranks of the pair
    x_T = torch.randn(b, 4, F - F1, 32, 32, generator=g).cuda()
    x0 = torch.randn(b, 4, F1, 32, 32, generator=g).cuda()
    c = torch.randn(b, F, 77, 768, generator=g).cuda()
    uc = torch.randn(b, 1, 77, 768, generator=g).expand(-1, F, -1, -1).contiguous().cuda()
    shape = (b, 4, F - F1, 32, 32)


    def run(sampler, reps=args.reps):
        out, best = None, 1e9
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ddim_sample_latents(sampler, net, shape, c, x_T, x0, ddim_steps=args.steps, scale=7.5, uc=uc)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t = torch.tensor([best], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # max over ranks, device-timed
        return out, float(t.item())


    single, ms_single = run(DDIMSampler(torch.device("cuda", local)))
    nccl, ms_nccl = run(DDIMSampler(torch.device("cuda", local)).enable_cfg_branch_split(group, branch, transport="nccl"))
    split, ms_split = run(DDIMSampler(torch.device("cuda", local)).enable_cfg_branch_split(group, branch, transport="p2p"))
    transports_agree = bool(torch.equal(nccl, split))

    same = bool(torch.equal(single, split))
    rel = float((single - split).norm() / single.norm())
    other = [torch.empty_like(split) for _ in range(world)]
    dist.all_gather(other, split)
    ranks_agree = all(bool(torch.equal(o, split)) for o in other)
    flags = torch.tensor([int(rel <= TOL), int(ranks_agree), int(same), int(transports_agree)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok_all = ok_all and bool(flags[0]) and bool(flags[1]) and bool(flags[3])
    if rank == 0:
        ne = args.steps + 1
        print(f"CFG-branch split, {b} clip(s) x {F} frames ({F1} ref), {ne} evaluations, guidance 7.5, {world} GPUs")
        print(f"  single GPU ([uc; c] batch of {2 * b}): {ms_single:8.1f} ms per pass = {ms_single / ne:.3f} ms per evaluation")
        print(f"  branch split, NCCL all-gather of eps + update kernel per step: {ms_nccl:8.1f} ms per pass = "
              f"{ms_nccl / ne:.3f} ms per step  -> latency x{ms_single / ms_nccl:.2f}")
        print(f"  branch split, exchange inside the update kernel over NVLink peer memory: {ms_split:8.1f} ms per pass = "
              f"{ms_split / ne:.3f} ms per step  -> latency x{ms_single / ms_split:.2f}")
        print(f"  p2p == nccl bit for bit: {bool(flags[3])}")
        print(f"  latents vs the single-GPU run: rel-L2 {rel:.3e} (<= {TOL:g}: {bool(flags[0])}; bit-identical: {bool(flags[2])}); "
              f"both ranks of the pair hold bit-identical latents: {bool(flags[1])}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
