"""CUDA-graph capture of one UNet evaluation (~700 kernel launches) for replay across the 31 DDIM steps.

The reference runs ~2000 eager op launches per forward and captures nothing; here the launch-bound
DAG is recorded once per (model, shapes, cond_frame) and replayed with static input buffers.  The text K/V of the
cross-attention layers live OUTSIDE the graph (they depend only on the context): they are recomputed in place,
eagerly, whenever a different context tensor is presented, and the graph reads the same buffers.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedUNet:
    def __init__(self, unet, x_in: torch.Tensor, t_in: torch.Tensor, c_in: torch.Tensor, cond_frame: int, warmup: int = 2,
                 cfg_shared: bool = False):
        self.unet = unet
        self.cfg_shared = cfg_shared
        self.x = x_in.detach().clone().float().contiguous()
        self.t = t_in.detach().clone()
        self.c = c_in.detach().clone().contiguous()
        self.cond_frame = cond_frame
        self.precision = unet.precision
        self.residual_stream = getattr(unet, "residual_stream", None)
        self._src = (c_in, c_in._version)
        self.device = x_in.device
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):                   # packs weights, fills the text K/V cache, sets func attributes
                    unet(self.x, self.t, self.c, cond_frame=cond_frame, cfg_shared_input=self.cfg_shared)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._capture(unet, cond_frame)

    def _capture(self, unet, cond_frame: int) -> None:
        self.kv = unet._kv                                # keep the K/V buffers the captured kernels point at alive
        # ... and the packed weights: the captured kernels and tensor maps hold raw pointers into them.  If the model's
        # parameters change later (load_state_dict / .to() / reset_parameters) the model drops ITS reference and bumps its
        # version; `matches` then fails, so this graph is never replayed over stale weights, and the memory cannot be
        # recycled under it while it is alive.
        self.weights_version = unet._weights_version
        self.packed = (unet._packed, unet._packed32)
        before = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = unet(self.x, self.t, self.c, cond_frame=cond_frame, cfg_shared_input=self.cfg_shared)
        self.launches_per_replay = ops.LAUNCHES - before

    def matches(self, unet, x_in, c_in, cond_frame, cfg_shared: bool = False) -> bool:
        return (self.unet is unet and self.cfg_shared == cfg_shared and tuple(self.x.shape) == tuple(x_in.shape) and tuple(self.c.shape) == tuple(c_in.shape)
                and self.cond_frame == cond_frame and self.precision == unet.precision
                and self.residual_stream == getattr(unet, "residual_stream", None)
                and self.weights_version == unet._weights_version and self.device == x_in.device)

    def __call__(self, x_in: torch.Tensor, t_in: torch.Tensor, c_in: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self.device):
            return self._replay(x_in, t_in, c_in)

    def _replay(self, x_in: torch.Tensor, t_in: torch.Tensor, c_in: torch.Tensor) -> torch.Tensor:
        if self._src[0] is not c_in or self._src[1] != c_in._version:
            self.c.copy_(c_in)
            self.unet.compute_context_kv(self.c, out=self.kv)
            self._src = (c_in, c_in._version)
        self.x.copy_(x_in)
        self.t.copy_(t_in)
        self.graph.replay()
        ops._count(self.launches_per_replay)
        return self.out
