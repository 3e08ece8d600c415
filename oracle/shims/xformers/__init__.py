"""Shim of xformers==0.0.13: memory_efficient_attention + LowerTriangularMask restated in
plain PyTorch (softmax(q k^T / sqrt(d) + bias) v, fp32 softmax).  Test infrastructure only."""
from . import ops  # noqa: F401
