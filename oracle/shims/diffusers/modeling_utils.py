"""ModelMixin plumbing (diffusers 0.10.2 restated): an nn.Module with `.dtype`/`.device`."""
import torch
from torch import nn


class ModelMixin(nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def enable_xformers_memory_efficient_attention(self):
        for m in self.modules():
            if hasattr(m, "_use_memory_efficient_attention_xformers"):
                m._use_memory_efficient_attention_xformers = True
