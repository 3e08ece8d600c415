"""FSTextTransformer on the seer_b200 kernels (-m gpu) against golden outputs of the unmodified reference and against the
oracle, in both precisions.  Tolerances: fp32 path 1e-4, bf16 path 2e-2 (north_star's per-step tolerances; the module ends
in a LayerNorm, so its output is O(1) per element)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import seer_oracle as so  # noqa: E402  (checker only)
from seervideoldm_b200 import FSTextTransformer, SeerUNet  # noqa: E402
from seervideoldm_b200.weights import random_fstext_state_dict  # noqa: E402

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def gen(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.fixture(scope="module")
def fstext():
    sd = random_fstext_state_dict(16, 2, seed=0)
    m = FSTextTransformer(num_frames=16, num_layers=2)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fstext_vs_reference_golden(fstext, golden_dir, precision):
    m, _ = fstext
    m.set_precision(precision)
    g = torch.load(os.path.join(golden_dir, "fstext.pt"), weights_only=False)
    for case in g["cases"]:
        ctx = gen(case["ctx_seed"], case["b"], 77, 768)
        m.set_numframe(case["num_frames"])
        y = m(ctx.cuda())
        assert y.shape == (case["b"], case["num_frames"], 77, 768) and y.dtype == torch.float32
        err = so.rel_l2(y[:, :, ::g["token_stride"], ::g["channel_stride"]].cpu(), case["y_sub"])
        print(f"FSText {precision} b={case['b']} F={case['num_frames']}: rel-L2 vs reference {err:.3e}")
        assert err < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fstext_vs_oracle_ragged(fstext, precision):
    """Shorter text (L = 20 < 77 tokens: the position table is sliced, :476) and an odd frame count."""
    m, sd = fstext
    m.set_precision(precision)
    m.set_numframe(5)
    ctx = gen(81, 2, 20, 768)
    ref = so.fstext_forward(sd, ctx, 5)
    y = m(ctx.cuda())
    err = so.rel_l2(y.cpu(), ref)
    print(f"FSText {precision} L=20 F=5: rel-L2 vs oracle {err:.3e}")
    assert err < TOL[precision]


def test_fstext_eight_layers_feeds_unet(fstext):
    """The shipped configuration (inference.py:88: num_frames 16, num_layers 8) end to end into the UNet's context input."""
    sd = random_fstext_state_dict(16, 8, seed=1)
    m = FSTextTransformer(num_frames=16, num_layers=8)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.set_numframe(4)
    ctx = gen(82, 1, 77, 768)
    y = m(ctx.cuda())
    ref = so.fstext_forward(sd, ctx, 4)
    err = so.rel_l2(y.cpu(), ref)
    print(f"FSText bf16, 8 layers: rel-L2 vs oracle {err:.3e}")
    assert err < 2e-2
    from seervideoldm_b200.config import sd15_config
    from seervideoldm_b200.weights import random_state_dict
    net = SeerUNet(sample_size=32, cross_attention_dim=768)
    net.load_state_dict(random_state_dict(sd15_config(sample_size=32), seed=0), strict=True)
    net = net.cuda().eval()
    x = gen(83, 1, 4, 4, 8, 8)
    t = torch.tensor([500])
    eps = net(x.cuda(), t.cuda(), y)                                  # per-frame sub-instruction embedding as context
    ref_eps = so.unet_forward(net.state_dict() and random_state_dict(sd15_config(sample_size=32), seed=0), x, t, ref, 0)
    assert so.rel_l2(eps.cpu(), ref_eps) < 2e-2
