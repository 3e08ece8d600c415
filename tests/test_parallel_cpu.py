"""world_size-2 gloo test of the clip sharding + all-gather (the N>1 path of bench.py / parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seervideoldm_b200.parallel import cfg_branch_group, gather_cfg_branches, gather_latents, sample_sharded, shard_clips


def test_shard_clips_partition():
    for n, w in [(64, 8), (7, 2), (5, 4), (3, 3)]:
        owned = [shard_clips(n, r, w) for r in range(w)]
        assert sorted(i for o in owned for i in o) == list(range(n))
        assert all(i % w == r for r, o in enumerate(owned) for i in o)


def _worker(rank, world, port, n_clips, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def sample_fn(ids):                                   # stand-in sampler: latents are a pure function of the clip id
        calls.append(list(ids))
        return torch.stack([torch.full((4, 3, 2, 2), float(i)) + torch.arange(4).reshape(4, 1, 1, 1) for i in ids])

    out = sample_sharded(sample_fn, n_clips, batch)
    q.put((rank, out, calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips,batch", [(6, 2), (5, 8)])
def test_sharded_sampling_gathers_in_clip_order(n_clips, batch):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([torch.full((4, 3, 2, 2), float(i)) + torch.arange(4).reshape(4, 1, 1, 1) for i in range(n_clips)])
    for rank, out, calls in res:
        assert torch.equal(out, want)                     # every rank holds all clips, in global order, bit-identical
        assert [i for c in calls for i in c] == shard_clips(n_clips, rank, 2)
        assert all(len(c) <= batch for c in calls)


def test_single_process_passthrough():
    x = torch.randn(3, 4)
    assert gather_latents(x, 3, 0, 1) is x


def _branch_worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    group, branch = cfg_branch_group()
    seen = []

    def sample_fn(ids):
        # stand-in for the DDIM loop: each rank "evaluates" only its CFG branch, the pair exchanges per step
        x = torch.stack([torch.full((4, 2, 2, 2), float(i)) for i in ids])
        for step in range(3):
            eps_local = x * (step + 1) + (100.0 if branch == 1 else 0.0)           # branch 1 = conditional
            eps = gather_cfg_branches(eps_local, group)
            e_uc, e_c = eps.chunk(2)                                                # the reference's [uc; c] order
            seen.append((torch.equal(e_uc, x * (step + 1)), torch.equal(e_c, x * (step + 1) + 100.0)))
            x = x + (e_c - e_uc) / 100.0                                            # both ranks apply the same update
        return x

    out = sample_sharded(sample_fn, n_clips, 8, cfg_branch_split=True)
    q.put((rank, branch, out, seen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clips", [(2, 1), (4, 3)])
def test_cfg_branch_split_pairs_exchange_in_uc_c_order(world, n_clips):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_branch_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([torch.full((4, 2, 2, 2), float(i) + 3.0) for i in range(n_clips)])
    for rank, branch, out, seen in res:
        assert branch == rank % 2
        assert torch.equal(out, want)                         # every rank: all clips, global order
        assert seen and all(a and b for a, b in seen)         # gathered eps is [uc; c] on both ranks of each pair


def test_cfg_branch_split_argument_checks():
    from seervideoldm_b200.ddim import DDIMSampler
    with pytest.raises(ValueError, match="even number"):
        cfg_branch_group(rank=0, world=3)
    s = DDIMSampler("cpu")
    with pytest.raises(ValueError, match="branch must be"):
        s.enable_cfg_branch_split(None, 2)
    assert s.enable_cfg_branch_split(None, 1)._branch == (None, 1) and s.disable_cfg_branch_split()._branch is None
    with pytest.raises(ValueError, match="transport"):
        s.enable_cfg_branch_split(None, 0, transport="mpi")
    assert s.enable_cfg_branch_split(None, 0)._transport == "nccl"                  # "auto" on a CPU sampler: the all-gather
    assert DDIMSampler(torch.device("cuda", 0)).enable_cfg_branch_split(None, 0)._transport == "p2p"
    from seervideoldm_b200.parallel import CfgPeerExchange
    with pytest.raises(ValueError, match="branch must be"):
        CfgPeerExchange(None, 3, 1024)


def _starved_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sample_sharded(lambda ids: torch.zeros(len(ids), 2), 1, 8)      # 1 clip, 2 ranks: rank 1 would own nothing
        q.put((rank, "no error"))
    except ValueError as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_fewer_clips_than_ranks_fails_on_every_rank():
    """n_clips < owners must raise on EVERY rank before any work (a rank-local error would leave the others in the all-gather)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_starved_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all("must be >= the number of owners" in res[r] for r in range(2)), res


def test_sampler_host_logic_cpu():
    """DDIMSampler host logic that needs no GPU: the graph cache is an LRU bounded by `max_graphs`, the reference's dead
    `null_cond_prob` branch raises NotImplementedError, the CFG prefix sharing can be switched off."""
    from seervideoldm_b200.ddim import DDIMSampler
    s = DDIMSampler("cpu", max_graphs=2, share_cfg_prefix=False)
    assert s.max_graphs == 2 and s.share_cfg_prefix is False and len(s._graphs) == 0
    s.make_schedule(30, verbose=False)
    x = torch.zeros(1, 4, 2, 8, 8)
    with pytest.raises(NotImplementedError, match="null_cond_prob"):
        s.p_sample_ddim(None, x, None, torch.tensor([1]), index=0, is_3d=True, null_cond_prob=0.1)
