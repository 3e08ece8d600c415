"""seervideoldm_b200 — B200-native implementation of Seer's DDIM+CFG denoising hot path.

Public surface mirrors the reference (seervideodiffusion/SeerVideoLDM):
    SeerUNet        <- seer/models/unet_3d_condition.py: SeerUNet
    FSTextTransformer <- seer/models/unet_3d_condition.py: FSTextTransformer
    DDIMSampler     <- ldm/models/diffusion/ddim_video.py: DDIMSampler
    ddim_sample     <- utils/ddim_sampling_utils.py: ddim_sample
    AutoencoderKL   <- diffusers 0.10.2 AutoencoderKL as the reference calls it (vae.decode / vae.encode, SD-1.5 VAE)
"""
from .config import UNetConfig, sd15_config  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require CUDA or the built library
    if name == "SeerUNet":
        from .unet import SeerUNet
        return SeerUNet
    if name == "FSTextTransformer":
        from .fstext import FSTextTransformer
        return FSTextTransformer
    if name == "DDIMSampler":
        from .ddim import DDIMSampler
        return DDIMSampler
    if name == "AutoencoderKL":
        from .vae import AutoencoderKL
        return AutoencoderKL
    if name in ("ddim_sample", "ddim_sample_latents"):
        from . import pipeline
        return getattr(pipeline, name)
    raise AttributeError(name)
