"""Import the UNMODIFIED reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there may call `load()`; it is used here to (1) pin `oracle/seer_oracle.py` against the real
reference and (2) generate the golden vectors under tests/golden/ (oracle/make_golden.py).

Two runtime patches, both documented in SURVEY.md §8c:
  * every CrossAttention gets `_use_memory_efficient_attention_xformers = True` (the other
    branch of seer/models/attention.py:681-694 crashes on broadcast — SURVEY F4);
  * `DDIMSampler.register_buffer` is replaced so buffers stay on the sampler's device instead
    of the hard-coded `.to("cuda")` (ldm/models/diffusion/ddim_video.py:21-25 — SURVEY F12).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SEER_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "seer", "models"))


def load() -> types.SimpleNamespace:
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import seer.models.attention as attention  # noqa: E402
    import seer.models.resnet as resnet  # noqa: E402
    import seer.models.unet_3d_blocks as blocks  # noqa: E402
    import seer.models.unet_3d_condition as cond  # noqa: E402
    import ldm.models.diffusion.ddim_video as ddim_video  # noqa: E402
    import ldm.modules.diffusionmodules.util as dutil  # noqa: E402

    def _register_buffer(self, name, attr):
        import torch
        if isinstance(attr, torch.Tensor):
            attr = attr.to(self.device)
        setattr(self, name, attr)

    ddim_video.DDIMSampler.register_buffer = _register_buffer
    return types.SimpleNamespace(attention=attention, resnet=resnet, blocks=blocks, cond=cond,
                                 ddim_video=ddim_video, dutil=dutil,
                                 SeerUNet=cond.SeerUNet, FSTextTransformer=cond.FSTextTransformer,
                                 DDIMSampler=ddim_video.DDIMSampler)


def enable_xformers_path(module) -> None:
    """Set the flag directly; the reference's own setter demands CUDA + real xformers
    (seer/models/attention.py:204-229)."""
    for m in module.modules():
        if hasattr(m, "_use_memory_efficient_attention_xformers"):
            m._use_memory_efficient_attention_xformers = True


def build_unet(ref, **cfg):
    """Reference SeerUNet with the SD-1.5 defaults of SURVEY.md §8c unless overridden."""
    kw = dict(sample_size=32, cross_attention_dim=768, attention_head_dim=8)
    kw.update(cfg)
    net = ref.SeerUNet(**kw).eval()
    enable_xformers_path(net)
    return net
