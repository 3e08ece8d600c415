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
LATENT = 32
DDIM_STEPS, EVALS, SCALE = 30, 31, 7.5
# Algorithmic FLOPs per batch-1 UNet evaluation (32x32 latents): SURVEY §8(d) / BASELINE.md §3 (causal-halved SCTA)
GFLOP_PER_EVAL = {12: 2887.5, 16: 3865.9}
# Work the implementation legitimately skips is subtracted, not credited, per batch-1 evaluation:
TEXT_KV_GFLOP = {12: 35.42, 16: 47.23}        # text K/V, cached across evaluations 2..31
# Upsample3D convs run as four 2x2-tap phase convs on the low-res image = 4/9 of the reference's 3x3-on-upsampled FLOPs
# (SURVEY App. A "upsample conv3x3" row: 203.84 GFLOP at F = 12, x 16/12 at F = 16): 5/9 of them are never executed
UPSAMPLE_SKIPPED_GFLOP = {12: 203.84 * 5 / 9, 16: 203.84 * 16 / 12 * 5 / 9}
# CFG shared prefix (SURVEY §8d "exploitable redundancy"): for the second half of every CFG pair the layers in front of the first
# cross-attention are not evaluated again — per SKIPPED batch-1 evaluation at F = 12 (x F/12; the self-attention core x (F/12) too,
# it is per frame): conv_in 0.28 + first ResNet 2 x 45.30 + proj_in 5.03 + q/k/v 15.10 + to_out 5.03 + cross q 5.03 +
# spatial self-attention core 16.11 GFLOP (SURVEY App. A level-0 rows divided by their instance counts)
CFG_PREFIX_GFLOP = {f: (0.28 + 2 * 45.30 + 5.03 + 15.10 + 5.03 + 5.03 + 16.11) * f / 12 for f in (12, 16)}

# BASELINE.json configs (name -> frames, reference frames, clips per GPU per sampling pass, total clips for strong scaling)
WORKLOADS = {
    "bridge": dict(frames=16, ref=1, clips=8, total=None,
                   desc="bridge (BASELINE.json configs[2]): 8 clips/GPU x (16f incl. 1 ref) 256^2 -> 32x32x4 latents"),
    "sthv2": dict(frames=12, ref=2, clips=1, total=None,
                  desc="sthv2 (BASELINE.json configs[1]): 1 clip/GPU x (12f incl. 2 ref) 256^2 -> 32x32x4 latents"),
    "sweep64": dict(frames=16, ref=1, clips=8, total=64,
                    desc="sweep64 (BASELINE.json configs[3]): 64 clips of 16f (1 ref) sharded by clip over the GPUs in local batches of 8"),
}
FRAMES, REF_FRAMES, CLIPS_PER_GPU = 16, 1, 8       # the default workload (bridge); run_ours rebinds them per --config


def set_workload(name: str):
    global FRAMES, REF_FRAMES, CLIPS_PER_GPU
    w = WORKLOADS[name]
    FRAMES, REF_FRAMES, CLIPS_PER_GPU = w["frames"], w["ref"], w["clips"]
    return w


def workload_name(name: str = "bridge") -> str:
    return f"{WORKLOADS[name]['desc']}, {EVALS}-eval DDIM, CFG {SCALE}"


def ncu_traffic():
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture of this round (tools/capture_gemm_full.sh -> profiles/r2_gemm_full.summary.json)."""
    path = os.path.join(ROOT, "profiles", "r2_gemm_full.summary.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        return j.get("traffic_bytes_per_launch_mean"), (f"mean DRAM bytes per launch over the {j.get('launches')} gemm_tc_kernel launches captured by "
                                                         f"tools/capture_gemm_full.sh ({j.get('what')}); algorithmic bytes of the same launches: "
                                                         f"{j.get('algorithmic_bytes_per_launch_mean'):.4g} (profiles/r2_gemm_full.summary.json)")
    return None, "no ncu --set full capture committed for this round's kernels (profiles/r2_gemm_full.summary.json missing)"


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
    set_workload(args.config)
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
            "config": {"workload": workload_name(args.config), "step": "bounded sample: " + cpu_sample_desc()},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": cpu_sample_desc()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
# config 5: attention stress (per-layer SCTA + FSText cross-attention at 64x64 latents, 16 frames, batch 4)
# --------------------------------------------------------------------------------------------------------------------
def run_stress(args):
    import torch
    from seervideoldm_b200 import ops
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = "cuda"
    B, F, heads = 4, 16, 8
    peaks = measured_peaks()
    rows, tot_ms, tot_fl = [], 0.0, 0.0
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    for h, C, reps in ((64, 320, 5), (32, 640, 5), (16, 1280, 5), (8, 1280, 1)):      # (level side, channels, layers at that level)
        d = C // heads
        M = B * F * h * h
        qkv = torch.randn(M, 3 * C, device=dev).bfloat16()
        q2 = torch.randn(M, C, device=dev).bfloat16()
        kv = torch.randn(B * F * 77, 2 * C, device=dev).bfloat16()
        ws = 0 if h <= 4 else (8 if h // 8 >= 4 else 4)
        L = F * (ws * ws if ws else h * h)
        nprob = B * heads * ((h // ws) ** 2 if ws else 1)
        cases = [("scta", lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], mode=ops.ATTN_SCTA, heads=heads, n_outer=B, F=F, H=h, W=h),
                  4.0 * L * (L + 1) / 2 * d * nprob),
                 ("cross", lambda: ops.attention(q2, kv[:, :C], kv[:, C:], mode=ops.ATTN_CROSS, heads=heads, n_outer=B * F, Lq=h * h, Lk=77),
                  4.0 * h * h * 77 * d * B * F * heads)]
        for name, fn, flops in cases:
            for _ in range(max(3, args.warmup)):
                fn()
            torch.cuda.synchronize()
            # the timed launches are replayed from a CUDA graph: the smallest of these kernels (50 us) are shorter than an eager
            # launch through the torch dispatcher, which would otherwise be what this loop measures
            n_launch = args.steps * 4
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(n_launch):
                    fn()
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n_launch
            rows.append({"op": f"{name} h={h} d={d}", "us": ms * 1e3, "tflops": flops / ms / 1e9, "layers": reps, "kernel": ops.last_attention_kernel()})
            tot_ms += ms * reps
            tot_fl += flops * reps
    clk = clocks.stop()
    tf = tot_fl / tot_ms / 1e9
    print(json.dumps({"metric": "attention TFLOP/s (SCTA + FSText cross-attention cores of one 64x64-latent forward, causal-halved FLOPs)",
                      "value": tf, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps * 4, "warmup": max(3, args.warmup),
                      "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                      "data": "synthetic", "config": {"workload": "stress (BASELINE.json configs[4]): SCTA + cross-attention at 64x64 latent, 16 frames, batch 4",
                                                      "l2": "q/k/v per launch 126-503 MB >= L2"},
                      "clocks": clk, "frac_of_sustained_tensor_peak": tf / peaks["sustained"], "per_launch": rows}))


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from seervideoldm_b200 import DDIMSampler, SeerUNet, ops
    from seervideoldm_b200.config import sd15_config
    from seervideoldm_b200.parallel import gather_latents, shard_clips
    from seervideoldm_b200.pipeline import ddim_sample_latents
    from seervideoldm_b200.weights import random_state_dict

    wl = set_workload(args.config)
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
    strong = wl["total"] is not None
    n_clips = wl["total"] if strong else b * world
    if strong and n_clips % (b * world):
        raise SystemExit(f"sweep64 needs {n_clips} clips to split into whole batches of {b} over {world} GPUs")
    clip_ids = shard_clips(n_clips, rank, world)          # round-robin ownership: clip i -> rank i mod world
    batches = [clip_ids[i: i + b] for i in range(0, len(clip_ids), b)]
    host = [[t.pin_memory() for t in batch_inputs(ids)] for ids in batches]
    resident = [[t.to(dev, non_blocking=True) for t in hb] for hb in host]
    shape = (b, 4, FRAMES - REF_FRAMES, LATENT, LATENT)
    out_host = torch.empty((n_clips,) + shape[1:], dtype=torch.float32).pin_memory()

    def sample(x_T, x0, c, uc):
        return ddim_sample_latents(sampler, net, shape, c, x_T, x0, ddim_steps=DDIM_STEPS, scale=SCALE, uc=uc)

    def one_pass_resident():
        lat = torch.cat([sample(*d) for d in resident]) if len(resident) > 1 else sample(*resident[0])
        return gather_latents(lat, n_clips, rank, world)

    def one_pass_e2e():
        outs = []
        for hb in host:
            d = [t.to(dev, non_blocking=True) for t in hb]          # H2D of this batch's inputs from pinned memory
            outs.append(sample(*d))
        allc = gather_latents(torch.cat(outs) if len(outs) > 1 else outs[0], n_clips, rank, world)
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
        last = one_pass_resident()
    clocks = ClockSampler(local)
    clocks.start()
    ms, launches = timed(one_pass_resident, args.steps)
    clk = clocks.stop()
    one_pass_e2e()
    ms_e2e, _ = timed(one_pass_e2e, args.steps)

    value = n_clips * args.steps / (ms * 1e-3)
    e2e_value = n_clips * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for hb in host for t in hb)
    d2h = out_host.numel() * out_host.element_size()

    # ---- --verify: sharding is bit-exact.  Rank 0 re-samples, on its own GPU and in the same local batch
    # composition, the clips another rank produced (its own when N = 1, through an eager sampler instead of the CUDA graph) and
    # requires torch.equal with the all-gathered result.
    verify = None
    if args.verify:
        if rank == 0:
            other = shard_clips(n_clips, (1 if world > 1 else 0), world)[:b]
            d = [t.to(dev) for t in batch_inputs(other)]
            eager = DDIMSampler(dev, use_cuda_graph=False)
            again = ddim_sample_latents(eager, net, shape, d[2], d[0], d[1], ddim_steps=DDIM_STEPS, scale=SCALE, uc=d[3])
            same = bool(torch.equal(again, last[other]))
            verify = {"clips": other, "produced_by_rank": 1 if world > 1 else 0, "recomputed_on_rank": 0, "bit_identical": same,
                      "max_abs_diff": float((again - last[other]).abs().max())}
            if not same:
                print(json.dumps({"verify": verify}), file=sys.stderr)
        barrier()

    # ---- roofline of the dominant kernel family (tcgen05 GEMM / implicit-GEMM conv), live CUDA events per launch ----
    peaks = measured_peaks()
    roof = None
    cpu_base = None
    attention = None
    if rank == 0:
        x_T, x0, c, uc = resident[0]
        x_in = torch.cat([torch.cat([x0, x_T], 2)] * 2)
        t_in, c_in = torch.full((2 * b,), 496, device=dev), torch.cat([uc, c])
        for _ in range(3):
            net(x_in, t_in, c_in)        # untimed eager passes: text K/V of the new context, allocator growth, clocks under load
        torch.cuda.synchronize()
        ops.PROFILE = []
        net(x_in, t_in, c_in)            # eager (no graph): CUDA events around each launch
        torch.cuda.synchronize()
        prof_all, ops.PROFILE = ops.PROFILE, None
        t_all_ms = sum(p[2].elapsed_time(p[3]) for p in prof_all)
        prof = [p for p in prof_all if p[0].startswith(("gemm ", "conv"))]       # the dominant kernel family only
        flops = sum(p[1] for p in prof)
        t_ms = sum(p[2].elapsed_time(p[3]) for p in prof)
        achieved = flops / (t_ms * 1e-3) / 1e12
        traffic, traffic_note = ncu_traffic()
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "gemm_tc_kernel<BN, CG> (tcgen05 GEMM + implicit-GEMM conv: 3x3, stride-2, 2x2 upsample phases)",
                "launches_timed": len(prof), "avg_launch_ms": t_ms / max(1, len(prof)),
                "flops_per_launch_avg": flops / max(1, len(prof)), "peak_source": peaks["source"] + ", sustained bf16",
                "frac_of_burst_peak": achieved / peaks["burst"],
                "share_of_step_time": t_ms / max(t_all_ms, 1e-9),
                "flops_note": "FLOPs actually executed by the timed launches (2 M N K each)"}
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
                fn()                                                    # warm-up (thread pool, allocator, caches)
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    fn()
                    ts.append(time.perf_counter() - t0)
            dt = statistics.median(ts)
            cpu_base = {"value": 1.0 / (EVALS * dt), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                        "sample": cpu_sample_desc() + "; median of 3 after 1 warm-up", "s_per_eval": dt, "nproc": os.cpu_count()}

    if rank == 0:
        evals_per_s = EVALS * args.steps * len(batches) / (ms * 1e-3)
        algo_tflop_per_clip = (EVALS * 2 * (GFLOP_PER_EVAL[FRAMES] - UPSAMPLE_SKIPPED_GFLOP[FRAMES])
                               - (EVALS - 1) * 2 * TEXT_KV_GFLOP[FRAMES] - EVALS * CFG_PREFIX_GFLOP[FRAMES]) / 1e3
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload_name(args.config),
                           "step": f"one {EVALS}-evaluation sampling pass over {len(clip_ids)} clips per GPU ({len(batches)} local batch(es) of {b})",
                           "global_batch_clips": n_clips, "unet_batch_per_gpu": 2 * b, "weights": "random-init SD-1.5-inflated SeerUNet, 1.083 B params",
                           "l2": "working set per evaluation (2.2 GB bf16 weights + multi-GB activations) >> 126 MB L2, no flush needed",
                           "ddim_evals_per_s": evals_per_s, "cuda_graph": True,
                           "algorithmic_tflop_per_clip": algo_tflop_per_clip,
                           "algorithmic_note": "reference-algorithm FLOPs minus the work legitimately skipped: cached text K/V (evaluations 2-31), "
                                               "5/9 of the Upsample3D conv FLOPs (2x2-tap phase convs on the low-res image) and the context-free "
                                               "front of the network for the second half of each CFG pair (conv_in, first ResNet, first text block up to to_q2)",
                           "achieved_tflops_whole_step": value * algo_tflop_per_clip,
                           "frac_of_sustained_peak_whole_step": value * algo_tflop_per_clip / world / peaks["sustained"]},
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "api": "seervideoldm_b200.pipeline.ddim_sample_latents (pinned host buffers)"},
                "gpu_launches": launches, "roofline": roof, "attention": attention}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if verify is not None:
            line["verify"] = verify
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if verify is not None and not verify["bit_identical"]:
        raise SystemExit("--verify: sharded result differs from the single-GPU recomputation")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="bridge", choices=sorted(WORKLOADS) + ["stress"],
                    help="BASELINE.json workload: bridge (default, the one the metric is quoted on), sthv2, sweep64 (64 clips, strong scaling), "
                         "stress (attention micro-benchmark at 64x64 latents)")
    ap.add_argument("--verify", action="store_true", help="rank 0 recomputes another rank's clips and requires bit-identity")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.config == "stress":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the attention stress config has no CPU arm; run --config bridge"}))
            return
        run_stress(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
