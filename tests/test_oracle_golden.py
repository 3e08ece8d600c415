"""Pin oracle/seer_oracle.py against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only.  Tolerance: fp32 run-to-run noise of the reference itself is
8.5e-7 rel-L2; the oracle must sit at that floor (<= 5e-6)."""
import os

import numpy as np
import pytest
import torch

from oracle import seer_oracle as so
from seervideoldm_b200.config import UNetConfig, sd15_config
from seervideoldm_b200.weights import random_state_dict

TOL = 5e-6


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.mark.parametrize("S,n", [(30, 31), (10, 10)])
def test_schedule_matches_reference(golden_dir, S, n):
    g = _load(golden_dir, f"schedule_{S}.pt")
    sch = so.make_schedule(S)
    assert len(sch.timesteps) == n
    assert np.array_equal(sch.timesteps, g["timesteps"].numpy())
    assert torch.equal(sch.alphas, g["alphas"])                                  # bit-exact fp32
    assert torch.equal(sch.alphas_prev, g["alphas_prev"].float())
    assert torch.equal(sch.sqrt_one_minus_alphas, g["sqrt_one_minus_alphas"])
    assert torch.equal(sch.alphas_cumprod, g["alphas_cumprod_fp32"])
    assert float(sch.sigmas.abs().max()) == 0.0


def test_schedule_known_answers():
    """SURVEY §8c known answers (1)-(2)."""
    sch = so.make_schedule(30)
    assert sch.timesteps[0] == 1 and sch.timesteps[30] == 991 and len(sch.timesteps) == 31
    assert abs(float(sch.alphas[30]) - 8.611506e-4) < 1e-9
    assert abs(float(sch.alphas[0]) - 0.99979734) < 1e-7
    assert abs(float(sch.alphas_prev[0]) - 0.99989998) < 1e-7
    assert abs(float(sch.alphas[15]) - 0.3389531672) < 1e-7


def test_scta_index_order_bit_exact(golden_dir):
    g = _load(golden_dir, "scta_index.pt")
    for key, seqs in g.items():
        if key.startswith("heads"):
            continue
        f, h, w = (int(v) for v in key.split("x"))
        assert np.array_equal(so.scta_sequences(f, h, w), seqs.numpy()), key
    t = torch.arange(2 * 3 * 16).reshape(2, 3, 16).float()
    mine = so._heads(t, 4).reshape(8, 3, 4)
    assert torch.equal(mine.long(), g["heads_to_batch_2x3x16_h4"])


def test_scta_causal_predicate_equals_reference_dependency(golden_dir):
    """closed form == autograd dependency pattern of the reference module."""
    g = _load(golden_dir, "scta_dependency.pt")
    for key, dep in g.items():
        f, h, w = (int(v) for v in key.split("x"))
        assert np.array_equal(so.scta_allowed(f, h, w), dep.numpy()), key


def test_window_size_rule():
    assert [so.scta_window_size(h) for h in (2, 4, 8, 16, 32, 64)] == [0, 0, 4, 4, 8, 8]


def test_module_level_golden(golden_dir):
    g = _load(golden_dir, "modules.pt")
    cfg = so.OracleCfg()
    for h in (4, 8, 16, 32):
        m = g[f"scta_h{h}"]
        y = so.scta(m["sd"], "", m["x"], 8)
        assert so.rel_l2(y, m["y"]) < TOL, h
    for cf in (0, 1, 2):
        m = g[f"temporal_xf_cf{cf}"]
        y = so.spatial_transformer(m["sd"], "", m["x"], None, True, cf, cfg)
        assert so.rel_l2(y, m["y"]) < TOL, cf
    m = g["text_xf"]
    y = so.spatial_transformer(m["sd"], "", m["x"], m["c"], False, 0, cfg)
    assert so.rel_l2(y, m["y"]) < TOL
    m = g["resnet_96_64"]
    y = so.resnet_block(m["sd"], "", m["x"], m["emb"], cfg)
    assert so.rel_l2(y, m["y"]) < TOL


def test_unet_narrow_golden(golden_dir):
    g = _load(golden_dir, "unet_narrow.pt")
    cfg = UNetConfig(sample_size=32, **g["cfg"])
    sd = random_state_dict(cfg, g["weight_seed"])
    ocfg = so.OracleCfg(block_out_channels=cfg.block_out_channels, cross_attention_dim=cfg.cross_attention_dim)
    for case in g["cases"]:
        y = so.unet_forward(sd, case["x"], case["t"], case["c"], case["cond_frame"], ocfg)
        assert so.rel_l2(y, case["y"]) < TOL


def test_ddim_loop_narrow_golden(golden_dir):
    g = _load(golden_dir, "ddim_loop_narrow.pt")
    cfg = UNetConfig(sample_size=32, **g["cfg"])
    sd = random_state_dict(cfg, g["weight_seed"])
    ocfg = so.OracleCfg(block_out_channels=cfg.block_out_channels, cross_attention_dim=cfg.cross_attention_dim)
    calls = []

    def unet_fn(x, t, c, cond_frame):
        calls.append((tuple(x.shape), t.tolist(), cond_frame))
        return so.unet_forward(sd, x, t, c, cond_frame, ocfg)

    lat, inter = so.ddim_sample_latents(unet_fn, g["xT"], g["c"], g["x0"], g["S"], g["scale"], g["uc"])
    assert calls == g["calls"]                       # 31 evaluations, [uc; c] batching, t = 991 ... 1
    assert len(calls) == 31 and calls[0][1] == [991, 991] and calls[-1][1] == [1, 1]
    assert so.rel_l2(lat, g["latents"]) < 5e-5       # 31 chained evaluations amplify the fp32 floor
    assert len(inter["x_inter"]) == len(g["x_inter"])
    for a, b in zip(inter["pred_x0"], g["pred_x0"]):
        assert so.rel_l2(a, b) < 5e-5


def test_unet_sd15_width_golden(golden_dir):
    g = _load(golden_dir, "unet_sd15.pt")
    cfg = sd15_config(sample_size=32)
    sd = random_state_dict(cfg, g["weight_seed"])
    gen = lambda seed, *shape: torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
    for case in g["cases"][:1]:
        B, Fr, H = case["shape"]
        x, c = gen(case["x_seed"], B, 4, Fr, H, H), gen(case["c_seed"], B, Fr, 77, 768)
        y = so.unet_forward(sd, x, case["t"], c, case["cond_frame"])
        assert so.rel_l2(y, case["y"]) < TOL


def test_ddim_update_is_reference_order():
    x, e = torch.randn(2, 4, 3, 8, 8), torch.randn(2, 4, 3, 8, 8)
    sch = so.make_schedule(30)
    f = lambda v: torch.full((2, 1, 1, 1, 1), float(v))
    xp, p0 = so.ddim_update(x, e, f(sch.alphas[7]), f(sch.alphas_prev[7]), f(sch.sqrt_one_minus_alphas[7]))
    a_t, a_p = sch.alphas[7], sch.alphas_prev[7]
    p0_ref = (x - torch.sqrt(1 - a_t) * e) / torch.sqrt(a_t)
    assert torch.equal(p0, p0_ref)
    assert torch.equal(xp, torch.sqrt(a_p) * p0_ref + torch.sqrt(1 - a_p) * e)


def test_fstext_oracle_vs_reference_golden(golden_dir):
    """oracle.fstext_forward against outputs of the unmodified reference FSTextTransformer (oracle/make_golden.py)."""
    from seervideoldm_b200.weights import fstext_schema, random_fstext_state_dict
    g = torch.load(os.path.join(golden_dir, "fstext.pt"), weights_only=False)
    sd = random_fstext_state_dict(g["num_frames"], g["num_layers"], seed=g["weight_seed"])
    assert list(sd.keys()) == list(fstext_schema(g["num_frames"], g["num_layers"]).keys()) and len(sd) == 72
    assert sum(v.numel() for k, v in sd.items() if not k.endswith("freqs")) == 55100160      # reference parameter count
    for case in g["cases"]:
        ctx = torch.randn(case["b"], 77, 768, generator=torch.Generator().manual_seed(case["ctx_seed"]))
        y = so.fstext_forward(sd, ctx, case["num_frames"])
        assert y.shape == (case["b"], case["num_frames"], 77, 768)
        assert so.rel_l2(y[:, :, ::g["token_stride"], ::g["channel_stride"]], case["y_sub"]) < 5e-6
        assert abs(float(y.mean()) - case["mean"]) < 1e-5 and abs(float(y.std()) - case["std"]) < 1e-5
