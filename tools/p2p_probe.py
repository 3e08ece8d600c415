#!/usr/bin/env python
"""Does torch symmetric memory (CUDA VMM peer mapping over NVLink) work on this box?  torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
t = symm_mem.empty(1024, dtype=torch.float32, device="cuda")
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
t.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
got = float(peer.sum().item())
hdl.barrier(channel=1)
print(f"rank {rank}: peer buffer sum {got} (expected {1024.0 * (((rank + 1) % world) + 1)}), multicast {hdl.has_multicast_support(torch.device('cuda').type, torch.cuda.current_device()) if hasattr(hdl, 'has_multicast_support') else '?'}")
dist.barrier()
dist.destroy_process_group()
