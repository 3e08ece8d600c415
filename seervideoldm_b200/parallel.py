"""Multi-GPU sampling: independent clips sharded across ranks, one all-gather of the final latents.

The reference's only parallelism is data parallel over the dataloader (inference.py:116-118) with an
`accelerator.gather` after sampling (utils/ddim_sampling_utils.py:9-12,60-63); there is no collective inside
the denoising step, so none is invented here: rank r owns clips r, r+W, r+2W, ... and the
only exchange is `all_gather_into_tensor` of (n_local, 4, F2, H, W) fp32 latents over NCCL/NVLink (gloo on CPU).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

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


def sample_sharded(sample_fn: Callable[[Sequence[int]], torch.Tensor], n_clips: int, batch: int) -> torch.Tensor:
    """Run `sample_fn(clip_ids) -> (len(ids), ...) latents` over this rank's clips in local batches, then all-gather."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_clips(n_clips, rank, world)
    outs = [sample_fn(mine[i: i + batch]) for i in range(0, len(mine), batch)]
    local = torch.cat(outs) if outs else None
    if local is None:
        raise ValueError("rank has no clips: n_clips must be >= world size")
    return gather_latents(local, n_clips, rank, world)
