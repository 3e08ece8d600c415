#!/usr/bin/env python
"""bench.py — headline benchmark of the Seer denoising hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): video clips/sec for 16-frame 256x256 clips, "30-step" (= 31-evaluation) DDIM
sampling under classifier-free guidance 7.5.  Workload = BASELINE.json configs[2] ("Bridge setting": batch 8,
16 frames / 1 reference frame, 32x32x4 latents, bf16 on 1 B200) — the configuration the metric is quoted on; it
fits one GPU.  One bench "step" = one full sampling pass (31 CFG evaluations of the 3-D UNet + fused DDIM
updates) over one local batch of 8 clips.  With N GPUs every rank samples its own 8 clips (weak scaling, no
collective inside the step) and the final latents are all-gathered once per pass (NCCL).

Synthetic data, random-init weights of the named architecture (1.08 B parameters, `proj_out` re-randomised so
the attention paths are live — SURVEY F9).  One JSON line is printed by rank 0.

`--impl reference` times the reference algorithm's CPU path (the unmodified reference when /root/reference is
present, else the oracle port) on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "video clips/sec (16f 256^2, 30-step DDIM+CFG)"
UNIT = "clips/s"
FRAMES, REF_FRAMES, LATENT, CLIPS_PER_GPU = 16, 1, 32, 8
DDIM_STEPS, EVALS, SCALE = 30, 31, 7.5
# Algorithmic FLOPs per batch-1 UNet evaluation at (F=16, 32x32): SURVEY §8(d) / BASELINE.md §3 (causal-halved SCTA)
GFLOP_PER_EVAL = 3865.9
TEXT_KV_GFLOP = 47.23          # cached across evaluations 2..31 -> subtracted, not credited
# dram__bytes_read.sum + dram__bytes_write.sum per gemm_tc_kernel launch from the committed `ncu --set full` capture
# (profiles/r1_gemm_full_v4.summary.txt: 8 level-1 launches, mean 418.5 MB against 454.0 MB algorithmic)
NCU_TRAFFIC = 418.5e6
NCU_TRAFFIC_NOTE = ("bytes per launch, mean of 8 level-1 GEMM launches of one evaluation under ncu --set full "
                    "(profiles/r1_gemm_full_v4.summary.txt); their algorithmic bytes: 454.0e6")


def workload_name() -> str:
    return (f"bridge: {CLIPS_PER_GPU} clips/GPU x ({FRAMES}f incl. {REF_FRAMES} ref) 256^2 -> {LATENT}x{LATENT}x4 latents, "
            f"{EVALS}-eval DDIM, CFG {SCALE}")


def clip_inputs(clip_id: int):
    """x_T, x0_emb, c, uc drawn in that order from Generator(1000 + clip_id)."""
    import torch
    g = torch.Generator().manual_seed(1000 + clip_id)
    f2 = FRAMES - REF_FRAMES
    x_T = torch.randn(1, 4, f2, LATENT, LATENT, generator=g)
    x0 = torch.randn(1, 4, REF_FRAMES, LATENT, LATENT, generator=g)
    c = torch.randn(1, FRAMES, 77, 768, generator=g)
    uc = torch.randn(1, 1, 77, 768, generator=g)
    return x_T, x0, c, uc


def batch_inputs(clip_ids):
    import torch
    parts = [clip_inputs(i) for i in clip_ids]
    x_T, x0, c, uc = (torch.cat([p[k] for p in parts]) for k in range(4))
    return x_T, x0, c, uc.expand(-1, FRAMES, -1, -1).contiguous()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, burst=1590.0, sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference algorithm on host cores, bounded sample
# --------------------------------------------------------------------------------------------------------------------
def cpu_eval_fn():
    """Returns (callable running ONE CFG evaluation of one 16-frame clip on CPU fp32, kind)."""
    import torch
    from oracle import reference_loader as rl, seer_oracle as so
    from seervideoldm_b200.config import sd15_config
    from seervideoldm_b200.weights import random_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    sd = random_state_dict(sd15_config(sample_size=32), seed=0)
    x_T, x0, c, uc = batch_inputs([0])
    x_in = torch.cat([torch.cat([x0, x_T], 2)] * 2)
    t_in = torch.full((2,), 991, dtype=torch.long)
    c_in = torch.cat([uc, c])
    if rl.available():
        ref = rl.load()
        net = rl.build_unet(ref)
        net.load_state_dict(sd, strict=True)
        with torch.no_grad():
            return (lambda: net(x_in, t_in, c_in, cond_frame=0)), "reference"
    return (lambda: so.unet_forward(sd, x_in, t_in, c_in, 0)), "port"


def cpu_sample_desc():
    return (f"1 CFG UNet evaluation (UNet batch 2 = 1 clip, {FRAMES}f, {LATENT}x{LATENT} latent, fp32); "
            f"clips/s = 1 / ({EVALS} x s_per_eval)")


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, kind = cpu_eval_fn()
    with torch.no_grad():
        for _ in range(max(0, args.warmup)):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt = (time.perf_counter() - t0) / args.steps
    v = 1.0 / (EVALS * dt)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(), "step": "bounded sample: " + cpu_sample_desc()},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": cpu_sample_desc()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from seervideoldm_b200 import DDIMSampler, SeerUNet, ops
    from seervideoldm_b200.config import sd15_config
    from seervideoldm_b200.parallel import gather_latents
    from seervideoldm_b200.pipeline import ddim_sample_latents
    from seervideoldm_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = sd15_config(sample_size=32)
    net = SeerUNet(sample_size=32, cross_attention_dim=768)
    net.load_state_dict(random_state_dict(cfg, seed=0), strict=True)
    net = net.to(dev).eval()
    sampler = DDIMSampler(dev)

    b = CLIPS_PER_GPU
    n_clips = b * world
    clip_ids = list(range(rank, n_clips, world))          # round-robin ownership (parallel.shard_clips)
    host = [t.pin_memory() for t in batch_inputs(clip_ids)]
    x_T, x0, c, uc = (t.to(dev, non_blocking=True) for t in host)
    shape = (b, 4, FRAMES - REF_FRAMES, LATENT, LATENT)
    out_host = torch.empty((n_clips,) + shape[1:], dtype=torch.float32).pin_memory()

    def one_pass_resident():
        lat = ddim_sample_latents(sampler, net, shape, c, x_T, x0, ddim_steps=DDIM_STEPS, scale=SCALE, uc=uc)
        return gather_latents(lat, n_clips, rank, world)

    def one_pass_e2e():
        d = [t.to(dev, non_blocking=True) for t in host]            # H2D of this pass's inputs from pinned memory
        lat = ddim_sample_latents(sampler, net, shape, d[2], d[0], d[1], ddim_steps=DDIM_STEPS, scale=SCALE, uc=d[3])
        allc = gather_latents(lat, n_clips, rank, world)
        out_host.copy_(allc, non_blocking=True)                     # D2H of the pass's result
        return allc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)              # max over ranks, device-timed
        return float(ms.item()), ops.LAUNCHES - l0

    for _ in range(max(3, args.warmup)):
        one_pass_resident()
    clocks = ClockSampler(local)
    clocks.start()
    ms, launches = timed(one_pass_resident, args.steps)
    clk = clocks.stop()
    one_pass_e2e()
    ms_e2e, _ = timed(one_pass_e2e, args.steps)

    value = n_clips * args.steps / (ms * 1e-3)
    e2e_value = n_clips * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = out_host.numel() * out_host.element_size()

    # ---- roofline of the dominant kernel family (tcgen05 GEMM / implicit-GEMM conv), live CUDA events per launch ----
    peaks = measured_peaks()
    roof = None
    cpu_base = None
    attention = None
    if rank == 0:
        x_in = torch.cat([torch.cat([x0, x_T], 2)] * 2)
        t_in, c_in = torch.full((2 * b,), 496, device=dev), torch.cat([uc, c])
        net(x_in, t_in, c_in)            # untimed eager pass: text K/V of the new context + allocator growth outside the graph pool
        torch.cuda.synchronize()
        ops.PROFILE = []
        net(x_in, t_in, c_in)            # eager (no graph): CUDA events around each launch
        torch.cuda.synchronize()
        prof_all, ops.PROFILE = ops.PROFILE, None
        t_all_ms = sum(p[2].elapsed_time(p[3]) for p in prof_all)
        prof = [p for p in prof_all if p[0].startswith(("gemm ", "conv3x3 "))]       # the dominant kernel family only
        flops = sum(p[1] for p in prof)
        t_ms = sum(p[2].elapsed_time(p[3]) for p in prof)
        achieved = flops / (t_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["sustained"], "traffic": NCU_TRAFFIC, "traffic_note": NCU_TRAFFIC_NOTE, "kernel": "gemm_tc_kernel<BN> (tcgen05 GEMM + implicit-GEMM conv3x3)",
                "launches_timed": len(prof), "avg_launch_ms": t_ms / max(1, len(prof)),
                "flops_per_launch_avg": flops / max(1, len(prof)), "peak_source": peaks["source"] + ", sustained bf16",
                "frac_of_burst_peak": achieved / peaks["burst"],
                "share_of_step_time": t_ms / max(t_all_ms, 1e-9)}
        # attention cores (SCTA / spatial / cross; BASELINE.json metric: "attn % of peak"), same live CUDA-event pass
        att = [p for p in prof_all if p[0].startswith("attention ")]
        att_ms = sum(p[2].elapsed_time(p[3]) for p in att)
        att_tf = sum(p[1] for p in att) / (att_ms * 1e-3) / 1e12 if att_ms > 0 else 0.0
        by_kind = {}
        for kind in ("scta", "spatial", "cross"):
            sel = [p for p in att if p[0].startswith("attention " + kind)]
            ms_k = sum(p[2].elapsed_time(p[3]) for p in sel)
            if ms_k > 0:
                by_kind[kind] = {"tflops": sum(p[1] for p in sel) / (ms_k * 1e-3) / 1e12, "ms_per_eval": ms_k, "launches": len(sel)}
        attention = {"achieved": att_tf, "unit": "TFLOP/s (causal-halved algorithmic FLOPs)", "peak": peaks["sustained"],
                     "frac_of_tensor_peak": att_tf / peaks["sustained"], "share_of_step_time": att_ms / max(t_all_ms, 1e-9),
                     "by_kind": by_kind,
                     "note": "d=40 level is MUFU(ex2)-bound: 160 tensor FLOP per exp2 caps it near 31 % of the tensor peak "
                             "(profiles/r1_attention_tc.summary.txt)"}
        if world == 1 and not args.no_cpu_baseline:
            fn, kind = cpu_eval_fn()
            with torch.no_grad():
                t0 = time.perf_counter()
                fn()
                dt = time.perf_counter() - t0
            cpu_base = {"value": 1.0 / (EVALS * dt), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                        "sample": cpu_sample_desc(), "s_per_eval": dt}

    if rank == 0:
        evals_per_s = EVALS * args.steps / (ms * 1e-3)
        algo_tflop_per_clip = (EVALS * 2 * GFLOP_PER_EVAL - (EVALS - 1) * 2 * TEXT_KV_GFLOP) / 1e3
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": workload_name(), "step": f"one {EVALS}-evaluation sampling pass over {b} clips per GPU",
                           "global_batch_clips": n_clips, "unet_batch_per_gpu": 2 * b, "weights": "random-init SD-1.5-inflated SeerUNet, 1.083 B params",
                           "l2": "working set per evaluation (2.2 GB bf16 weights + multi-GB activations) >> 126 MB L2, no flush needed",
                           "ddim_evals_per_s": evals_per_s, "cuda_graph": True,
                           "algorithmic_tflop_per_clip": algo_tflop_per_clip,
                           "achieved_tflops_whole_step": value * algo_tflop_per_clip,
                           "frac_of_sustained_peak_whole_step": value * algo_tflop_per_clip / world / peaks["sustained"]},
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "api": "seervideoldm_b200.pipeline.ddim_sample_latents (pinned host buffers)"},
                "gpu_launches": launches, "roofline": roof, "attention": attention}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
