"""Multi-GPU sampling: independent clips sharded across ranks, one all-gather of the final latents.

The reference's only parallelism is data parallel over the dataloader (inference.py:116-118) with an
`accelerator.gather` after sampling (utils/ddim_sampling_utils.py:9-12,60-63); there is no collective inside
the denoising step, so none is invented here: rank r owns clips r, r+W, r+2W, ... and the
only exchange is `all_gather_into_tensor` of (n_local, 4, F2, H, W) fp32 latents over NCCL/NVLink (gloo on CPU).

Optional latency mode — CFG-branch split (BASELINE.json north_star "sharding independent clips and CFG branches",
SURVEY §8e row 2): when there are more GPUs than clips, ranks pair up (2p, 2p+1); the even rank evaluates the
unconditional branch, the odd rank the conditional one (UNet batch b instead of 2b), and the two exchange the noise
prediction once per DDIM step before both apply the CFG+DDIM update redundantly.  This is the one place data crosses GPUs
inside the step, so it is off by default.  Two transports:
  * "p2p" (default on CUDA): `CfgPeerExchange` — the update kernel itself pushes this rank's branch into the partner's
    receive slot through NVLink peer memory (symmetric memory), signals, waits for the partner and combines: one launch per
    step, no NCCL call in the loop (csrc/elementwise.cu `cfg_ddim_p2p_kernel`);
  * "nccl": `gather_cfg_branches`, one all-gather of (b,4,F,H,W) fp32 inside the pair (262 KB per clip at the bench shape),
    then the ordinary update kernel (also what the gloo CPU tests exercise).
Both ranks of a pair end with bit-identical latents, the two transports agree bit for bit, and since the LayerNorm
row-statistic producers use a batch-invariant tile plan the result is also bit-identical to the single-GPU `[uc; c]` batch
(`profiles/r2_cfg_branch_split_2gpu.txt`; latency x1.35 at 16 frames, x1.30 at 12 frames: the UNet at batch 1 is
launch-latency bound, not the exchange).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: clip i -> rank i mod world."""
    return list(range(rank, n_clips, world))


def gather_latents(local: torch.Tensor, n_clips: int, rank: int, world: int) -> torch.Tensor:
    """local: (n_local, ...) latents of this rank's clips (in shard_clips order) -> (n_clips, ...) on every rank,
    restored to global clip order.  Ranks with fewer clips are padded for the collective."""
    if world == 1:
        return local
    per = (n_clips + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.reshape(world, per, *local.shape[1:])
    res = torch.empty((n_clips,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = shard_clips(n_clips, r, world)
        res[ids] = out[r, : len(ids)]
    return res


def sample_sharded(sample_fn: Callable[[Sequence[int]], torch.Tensor], n_clips: int, batch: int,
                   cfg_branch_split: bool = False) -> torch.Tensor:
    """Run `sample_fn(clip_ids) -> (len(ids), ...) latents` over this rank's clips in local batches, then all-gather.
    With `cfg_branch_split` the unit of ownership is the rank PAIR (both ranks of a pair run `sample_fn` on the same
    clips, their sampler exchanging CFG branches per step) and the even rank's copy is the one kept."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if cfg_branch_split:
        if world % 2:
            raise ValueError("cfg_branch_split needs an even number of ranks")
        owner, owners = rank // 2, world // 2
    else:
        owner, owners = rank, world
    if n_clips < owners:          # every rank knows both numbers: all of them raise, none is left waiting in the collective
        raise ValueError(f"n_clips ({n_clips}) must be >= the number of owners ({owners}: ranks, or rank pairs)")
    mine = shard_clips(n_clips, owner, owners)
    outs = [sample_fn(mine[i: i + batch]) for i in range(0, len(mine), batch)]
    local = torch.cat(outs) if outs else None
    if local is None:
        raise ValueError("rank has no clips: n_clips must be >= the number of owners (ranks, or rank pairs)")
    if not cfg_branch_split:
        return gather_latents(local, n_clips, rank, world)
    per = (n_clips + owners - 1) // owners
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.reshape(world, per, *local.shape[1:])
    res = torch.empty((n_clips,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for o in range(owners):
        ids = shard_clips(n_clips, o, owners)
        res[ids] = out[2 * o, : len(ids)]
    return res


# ---------------------------------------------------------------------------------------------------------------------
# CFG-branch split (optional latency mode)
# ---------------------------------------------------------------------------------------------------------------------
def cfg_branch_group(rank: Optional[int] = None, world: Optional[int] = None) -> Tuple["dist.ProcessGroup", int]:
    """Pair ranks (2p, 2p+1) -> (this rank's pair group, branch index): 0 = unconditional, 1 = conditional — the
    order of the reference's `[uc; c]` batch (ddim_video.py:199-203).  Collective: every rank must call it."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world % 2:
        raise ValueError("CFG-branch split needs an even number of ranks")
    mine = None
    for p in range(world // 2):
        g = dist.new_group(ranks=[2 * p, 2 * p + 1])         # new_group is collective over the whole world
        if rank // 2 == p:
            mine = g
    return mine, rank % 2


def gather_cfg_branches(eps_local: torch.Tensor, group) -> torch.Tensor:
    """(b, C, F, H, W) noise prediction of this rank's branch -> (2b, C, F, H, W) in `[uc; c]` order on both ranks of
    the pair (all-gather concatenates along dim 0 in group-rank order = branch order)."""
    eps_local = eps_local.contiguous()
    out = torch.empty((2 * eps_local.shape[0],) + tuple(eps_local.shape[1:]), dtype=eps_local.dtype, device=eps_local.device)
    dist.all_gather_into_tensor(out, eps_local, group=group)
    return out


class CfgPeerExchange:
    """Per-step CFG-branch exchange through NVLink peer memory instead of an NCCL all-gather: the transport of
    `ops.cfg_ddim_update_p2p` (csrc/elementwise.cu `cfg_ddim_p2p_kernel`).

    One symmetric allocation per rank of the pair (`torch.distributed._symmetric_memory`: CUDA VMM memory every rank of the
    group maps into its own address space) holds two receive slots of `numel` fp32 noise-prediction values and two flag
    words; the partner's copy is mapped with `get_buffer`.  The sampler's update kernel pushes this rank's branch straight
    into the partner's slot, signals, waits for the partner's signal and combines — one launch per step, no NCCL call and
    no extra pass over eps inside the denoising loop.  Slots and flags alternate by step parity (see the kernel's header
    comment for why two are enough); `seq` increases by one per step for the life of the object.

    Collective over `group` (exactly two ranks): construct it on both ranks with the same `numel`."""

    FLAG_WORDS = 32             # the two flag words sit 64 B apart at the end of the allocation

    def __init__(self, group, branch: int, numel: int, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        if branch not in (0, 1):
            raise ValueError("branch must be 0 (unconditional) or 1 (conditional)")
        if dist.get_world_size(group) != 2:
            raise ValueError("CfgPeerExchange pairs exactly two ranks")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("CfgPeerExchange needs CUDA peer memory (NVLink / NVSwitch); on CPU use gather_cfg_branches")
        self.branch, self.group = int(branch), group
        self.numel = int((numel + 31) // 32 * 32)               # slots stay 128-byte aligned
        words = 2 * self.numel + self.FLAG_WORDS
        self._buf = symm_mem.empty(words, dtype=torch.float32, device=device)
        self._buf.zero_()
        self._hdl = symm_mem.rendezvous(self._buf, group)
        me = self._hdl.rank
        peer = self._hdl.get_buffer(1 - me, (words,), torch.float32)
        self._peer = peer
        self._local_recv = [self._buf[s * self.numel:(s + 1) * self.numel] for s in (0, 1)]
        self._peer_recv = [peer[s * self.numel:(s + 1) * self.numel] for s in (0, 1)]
        lf = self._buf[2 * self.numel:].view(torch.int32)
        pf = peer[2 * self.numel:].view(torch.int32)
        self._local_flag = [lf[0:1], lf[16:17]]
        self._peer_flag = [pf[0:1], pf[16:17]]
        self._counter = torch.zeros(1, dtype=torch.int32, device=device)
        self._seq = 0
        torch.cuda.synchronize(device)
        self._hdl.barrier(channel=0)                             # both allocations zeroed before either rank's first push

    def update(self, eps_local: torch.Tensor, x: torch.Tensor, cond_f: int, scale: float, coef) -> Tuple[torch.Tensor, torch.Tensor]:
        """eps_local (b, C, cond_f+F2, H, W): this rank's branch; x (b, C, F2, H, W); coef = the step's
        (sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), dir) -> (x_prev, pred_x0), identical on both ranks."""
        from . import ops
        if x.numel() > self.numel:
            raise ValueError(f"exchange sized for {self.numel} values per step, got {x.numel()}")
        self._seq += 1
        s = self._seq & 1
        return ops.cfg_ddim_update_p2p(eps_local, self.branch, self._peer_recv[s], self._local_recv[s], self._peer_flag[s],
                                       self._local_flag[s], self._counter, self._seq, x, cond_f, scale, float(coef[0]),
                                       float(coef[1]), float(coef[2]), float(coef[3]))
