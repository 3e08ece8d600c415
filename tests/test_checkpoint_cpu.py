"""Checkpoint / wire formats: diffusers-layout directories, accelerate `pytorch_model*.bin` files and
the safetensors restatement — all on CPU (loading never needs the CUDA library)."""
import json
import os
import struct
import warnings

import pytest
import torch

from seervideoldm_b200 import checkpoint as ck
from seervideoldm_b200.config import UNetConfig
from seervideoldm_b200.weights import random_fstext_state_dict, random_state_dict

NARROW = dict(sample_size=32, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64)


def _unet():
    from seervideoldm_b200.unet import SeerUNet
    return SeerUNet(**NARROW)


def test_safetensors_round_trip_all_dtypes(tmp_path):
    g = torch.Generator().manual_seed(0)
    t = {"f32": torch.randn(3, 5, generator=g), "f16": torch.randn(7, generator=g).half(),
         "bf16": torch.randn(2, 2, 2, generator=g).bfloat16(), "f64": torch.randn(4, generator=g).double(),
         "i64": torch.arange(6).reshape(2, 3), "u8": torch.arange(5, dtype=torch.uint8), "flag": torch.tensor([True, False]),
         "scalar": torch.tensor(3.5), "empty": torch.empty(0, 4)}
    p = str(tmp_path / "t.safetensors")
    ck.write_safetensors(p, t, metadata={"format": "pt"})
    back = ck.read_safetensors(p)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and back[k].shape == t[k].shape and torch.equal(back[k], t[k]), k
    # header is 8-byte aligned and the byte count in the file matches the header's offsets
    with open(p, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        hdr = json.loads(f.read(n))
    assert n % 8 == 0 and hdr["__metadata__"] == {"format": "pt"}
    assert os.path.getsize(p) == 8 + n + max(v["data_offsets"][1] for k, v in hdr.items() if k != "__metadata__")


def test_safetensors_agrees_with_the_published_library(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    g = torch.Generator().manual_seed(1)
    t = {"a.weight": torch.randn(8, 3, 3, 3, generator=g), "a.bias": torch.randn(8, generator=g).half(),
         "z": torch.randn(5, 2, generator=g).bfloat16(), "idx": torch.arange(4)}
    ours, theirs = str(tmp_path / "ours.safetensors"), str(tmp_path / "theirs.safetensors")
    ck.write_safetensors(ours, t)
    st.save_file(t, theirs)
    a, b = st.load_file(ours), ck.read_safetensors(theirs)            # each reader on the other writer's file
    for k in t:
        assert torch.equal(a[k], t[k]) and a[k].dtype == t[k].dtype
        assert torch.equal(b[k], t[k]) and b[k].dtype == t[k].dtype


def test_safetensors_rejects_corrupt_files(tmp_path):
    p = str(tmp_path / "bad.safetensors")
    with open(p, "wb") as f:
        f.write(b"\x01\x02")
    with pytest.raises(ValueError, match="8-byte"):
        ck.read_safetensors(p)
    with open(p, "wb") as f:
        f.write(struct.pack("<Q", 1 << 40) + b"{}")
    with pytest.raises(ValueError, match="header length"):
        ck.read_safetensors(p)
    hdr = json.dumps({"w": {"dtype": "F32", "shape": [4], "data_offsets": [0, 16]}}).encode()
    with open(p, "wb") as f:
        f.write(struct.pack("<Q", len(hdr)) + hdr + b"\0" * 8)            # data shorter than the header claims
    with pytest.raises(ValueError, match="spans bytes"):
        ck.read_safetensors(p)
    hdr = json.dumps({"w": {"dtype": "F8_E4M3", "shape": [4], "data_offsets": [0, 4]}}).encode()
    with open(p, "wb") as f:
        f.write(struct.pack("<Q", len(hdr)) + hdr + b"\0" * 4)
    with pytest.raises(ValueError, match="unsupported dtype"):
        ck.read_safetensors(p)


@pytest.mark.parametrize("safe", [False, True])
def test_save_pretrained_from_pretrained_round_trip(tmp_path, safe):
    from seervideoldm_b200.unet import SeerUNet
    net = _unet()
    sd = random_state_dict(UNetConfig(**NARROW), seed=5)
    net.load_state_dict(sd, strict=True)
    root = tmp_path / "sd15"
    net.save_pretrained(str(root / "unet"), safe_serialization=safe)
    cfg = json.load(open(root / "unet" / "config.json"))
    assert cfg["_class_name"] == "SeerUNet" and cfg["block_out_channels"] == [64, 128, 128, 128]
    # the reference's call, hub-only arguments included (inference.py:82-87)
    with warnings.catch_warnings():
        warnings.simplefilter("error")                                   # a complete checkpoint loads silently
        back = SeerUNet.from_pretrained(str(root), subfolder="unet", revision=None, low_cpu_mem_usage=False)
    assert not back.training and back.config.cross_attention_dim == 64 and back.config["sample_size"] == 32
    got = back.state_dict()
    assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)


def test_from_pretrained_inflates_a_2d_stable_diffusion_unet(tmp_path):
    """SD-1.5's UNet2DConditionModel file holds the non-temporal subset of the schema under the same names; extra config
    keys of newer diffusers versions and 2-D block-type names are ignored (the reference hard-codes the 3-D blocks)."""
    from seervideoldm_b200.unet import SeerUNet
    sd = random_state_dict(UNetConfig(**NARROW), seed=6)
    sd2d = {k: v.half() for k, v in sd.items() if "temporal_attentions." not in k}          # an fp16 hub variant
    d = tmp_path / "unet"
    os.makedirs(d)
    cfg = dict(_class_name="UNet2DConditionModel", _diffusers_version="0.6.0", act_fn="silu", attention_head_dim=8,
               block_out_channels=[64, 128, 128, 128], center_input_sample=False, cross_attention_dim=64,
               down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], downsample_padding=1, flip_sin_to_cos=True,
               freq_shift=0, in_channels=4, layers_per_block=2, mid_block_scale_factor=1, norm_eps=1e-05, norm_num_groups=32,
               out_channels=4, sample_size=32, up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3,
               use_linear_projection=False, upcast_attention=False)
    json.dump(cfg, open(d / "config.json", "w"))
    ck.write_safetensors(str(d / "diffusion_pytorch_model.fp16.safetensors"), sd2d)
    with pytest.warns(UserWarning, match="temporal-attention entries"):
        net, info = SeerUNet.from_pretrained(str(tmp_path), subfolder="unet", output_loading_info=True)
    n_temporal = sum("temporal_attentions." in k for k in sd)
    assert len(info["missing_keys"]) == n_temporal and all("temporal_attentions." in k for k in info["missing_keys"])
    assert info["unexpected_keys"] == []
    got = net.state_dict()
    assert all(v.dtype == torch.float32 for v in got.values())                                # upcast, parameters stay fp32
    assert all(torch.equal(got[k], sd2d[k].float()) for k in sd2d)
    # the temporal branch is as constructed: zero proj_out (attention.py:126-127), closed-form rotary table
    assert float(got["mid_block.temporal_attentions.0.proj_out.weight"].abs().max()) == 0.0
    assert torch.equal(got["mid_block.temporal_attentions.0.transformer_blocks.0.attn1.rotary_emb.freqs"],
                       sd["mid_block.temporal_attentions.0.transformer_blocks.0.attn1.rotary_emb.freqs"])


def test_from_pretrained_refuses_damaged_checkpoints(tmp_path):
    from seervideoldm_b200.unet import SeerUNet
    net = _unet()
    net.save_pretrained(str(tmp_path / "a"))
    sd = ck.read_state_dict(str(tmp_path / "a" / "diffusion_pytorch_model.bin"))
    # a spatial key missing: not an inflation, a truncated file
    cut = dict(sd); cut.pop("down_blocks.0.resnets.0.conv1.weight")
    torch.save(cut, tmp_path / "a" / "diffusion_pytorch_model.bin")
    with pytest.raises(RuntimeError, match="truncated"):
        SeerUNet.from_pretrained(str(tmp_path / "a"))
    # a wrong shape always raises
    wrong = dict(sd); wrong["conv_in.weight"] = torch.zeros(64, 8, 3, 3)
    torch.save(wrong, tmp_path / "a" / "diffusion_pytorch_model.bin")
    with pytest.raises(RuntimeError, match="size mismatch for conv_in.weight"):
        SeerUNet.from_pretrained(str(tmp_path / "a"))
    # unexpected keys are reported, not fatal
    extra = dict(sd); extra["class_embedding.weight"] = torch.zeros(3)
    torch.save(extra, tmp_path / "a" / "diffusion_pytorch_model.bin")
    with pytest.warns(UserWarning, match="were not used"):
        SeerUNet.from_pretrained(str(tmp_path / "a"))
    os.remove(tmp_path / "a" / "diffusion_pytorch_model.bin")
    with pytest.raises(OSError, match="no weights file"):
        SeerUNet.from_pretrained(str(tmp_path / "a"))
    with pytest.raises(OSError, match="not a directory"):
        SeerUNet.from_pretrained("runwayml/stable-diffusion-v1-5", subfolder="unet")


def test_load_seer_checkpoint_mirrors_the_reference_restore(tmp_path):
    """inference.py:119-128: pytorch_model.bin -> SeerUNet, pytorch_model_1.bin -> FSTextTransformer, strict; a DDP
    `module.` prefix is tolerated."""
    from seervideoldm_b200.fstext import FSTextTransformer
    net, fs = _unet(), FSTextTransformer(num_frames=16, num_layers=2)
    sd = random_state_dict(UNetConfig(**NARROW), seed=7)
    fsd = random_fstext_state_dict(16, 2, seed=7)
    d = tmp_path / "learned_sdunet-steps-200000"
    os.makedirs(d)
    torch.save({"module." + k: v for k, v in sd.items()}, d / "pytorch_model.bin")
    torch.save(fsd, d / "pytorch_model_1.bin")
    ck.load_seer_checkpoint(str(d), net, fs)
    assert all(torch.equal(v, sd[k]) for k, v in net.state_dict().items())
    assert all(torch.equal(v, fsd[k]) for k, v in fs.state_dict().items())
    part = dict(fsd); part.pop("norm.weight")
    torch.save(part, d / "pytorch_model_1.bin")
    with pytest.raises(RuntimeError, match="norm.weight"):
        ck.load_seer_checkpoint(str(d), None, fs)
    os.remove(d / "pytorch_model.bin")
    with pytest.raises(FileNotFoundError):
        ck.load_seer_checkpoint(str(d), net, None)
