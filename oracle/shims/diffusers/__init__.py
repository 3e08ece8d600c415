"""Shim of diffusers==0.10.2 (only what seer/models imports). Test infrastructure only."""
