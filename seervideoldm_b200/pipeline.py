"""`ddim_sample` — the pipeline entry the reference's scripts call
(/root/reference/utils/ddim_sampling_utils.py:20-42; callers inference.py:199, eval.py:222, inference_img.py:181).

`ddim_sample_latents` is the same call stopping before the VAE decode (the SD-1.5 AutoencoderKL is an external
pretrained network outside the denoising loop — out of scope)."""
from __future__ import annotations

import torch


@torch.no_grad()
def ddim_sample_latents(sampler, unet, shape, c, start_code, x0_emb, ddim_steps=10, scale=1.0, uc=None):
    if scale == 1.0:
        uc = None
    samples, _ = sampler.sample(unet=unet, S=ddim_steps, conditioning=c, batch_size=shape[0], shape=shape[1:], x0_emb=x0_emb,
                                verbose=False, unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0,
                                x_T=start_code, is_3d=True)
    return samples


@torch.no_grad()
def decode_latents(vae, samples):
    """Second half of the reference's `ddim_sample` (ddim_sampling_utils.py:37-41): `(n c f h w) -> ((n f) c h w)`, divide by the
    SD latent scale 0.18215, `vae.decode(z).sample`, back to `(n c f h w)`, map [-1, 1] -> [0, 1] and clamp.  `vae` is the
    caller's AutoencoderKL (any object with `.decode(z).sample`)."""
    n, ch, f, h, w = samples.shape
    z = 1 / 0.18215 * samples.permute(0, 2, 1, 3, 4).reshape(n * f, ch, h, w)
    x = vae.decode(z).sample
    x = x.reshape(n, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)
    return torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0)


@torch.no_grad()
def ddim_sample(sampler, unet, vae, shape, c, start_code, x0_emb, ddim_steps=10, scale=1.0, uc=None):
    samples = ddim_sample_latents(sampler, unet, shape, c, start_code, x0_emb, ddim_steps=ddim_steps, scale=scale, uc=uc)
    return decode_latents(vae, samples)
