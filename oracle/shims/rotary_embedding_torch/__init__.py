"""rotary-embedding-torch==0.1.5 restated (freqs_for='lang', theta=10000, interleaved pairs).
Call sites: /root/reference/seer/models/attention.py:480,529-530,650-651.  Test infrastructure only."""
import torch
from torch import nn


def rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def apply_rotary_emb(freqs, t, start_index=0):
    freqs = freqs.to(t)
    rot_dim = freqs.shape[-1]
    end_index = start_index + rot_dim
    t_left, t_mid, t_right = t[..., :start_index], t[..., start_index:end_index], t[..., end_index:]
    t_mid = (t_mid * freqs.cos()) + (rotate_half(t_mid) * freqs.sin())
    return torch.cat((t_left, t_mid, t_right), dim=-1)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.register_buffer("freqs", freqs)

    def forward(self, t):
        freqs = self.freqs
        freqs = torch.einsum("..., f -> ... f", t.type(freqs.dtype), freqs)
        return freqs.repeat_interleave(2, dim=-1)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        seq_len = t.shape[seq_dim]
        freqs = self.forward(torch.arange(seq_len, device=t.device))
        return apply_rotary_emb(freqs, t)
