"""world_size-2 gloo test of the clip sharding + all-gather (the N>1 path of bench.py / parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seervideoldm_b200.parallel import gather_latents, sample_sharded, shard_clips


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
