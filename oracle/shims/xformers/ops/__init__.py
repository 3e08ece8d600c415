"""xformers.ops (0.0.13) restated: 3-D (B', L, d) attention, optional causal bias.
Call site: /root/reference/seer/models/attention.py:622-630."""
import torch


class LowerTriangularMask:
    pass


def memory_efficient_attention(query, key, value, attn_bias=None, p=0.0, scale=None):
    d = query.shape[-1]
    scale = d ** -0.5 if scale is None else scale
    s = torch.baddbmm(
        torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device),
        query, key.transpose(-1, -2), beta=0, alpha=scale).float()
    if isinstance(attn_bias, LowerTriangularMask):
        lq, lk = s.shape[-2:]
        keep = torch.ones(lq, lk, dtype=torch.bool, device=s.device).tril()
        s = s.masked_fill(~keep, float("-inf"))
    elif attn_bias is not None:
        s = s + attn_bias
    p_ = s.softmax(dim=-1).to(value.dtype)
    return torch.bmm(p_, value)
