"""seervideoldm_b200 — B200-native implementation of Seer's DDIM+CFG denoising hot path."""
from .config import UNetConfig, sd15_config  # noqa: F401
