"""Stub: utils/ddim_sampling_utils.py imports imageio at module scope; nothing on the path uses it."""
