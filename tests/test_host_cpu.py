"""CPU-side tests (-m "not gpu"): host logic, state-dict schema, sampler schedule vs golden, the C-ABI library's
exports, and the rule that the product never imports the oracle."""
import ast
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from seervideoldm_b200 import _lib, packing
from seervideoldm_b200.config import UNetConfig, sd15_config
from seervideoldm_b200.weights import random_state_dict, unet_schema

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schema_known_answers():
    """SURVEY §8c (3),(8): 1006 entries, 1,082,772,804 parameters, 16 rotary buffers of shape (16,)."""
    s = unet_schema(sd15_config(sample_size=32))
    assert len(s) == 1006
    freqs = [k for k in s if k.endswith("rotary_emb.freqs")]
    assert len(freqs) == 16 and all(s[k] == (16,) for k in freqs)
    assert sum(int(np.prod(v)) for k, v in s.items() if k not in freqs) == 1_082_772_804
    assert len([k for k in s if k.endswith("conv_shortcut.weight")]) == 14
    assert len({k.rsplit(".resnets.", 1)[0] + k.rsplit(".resnets.", 1)[1][0] for k in s if ".resnets." in k}) == 22


def test_random_state_dict_is_deterministic_and_loads():
    from seervideoldm_b200.unet import SeerUNet
    cfg = UNetConfig(sample_size=32, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64)
    a, b = random_state_dict(cfg, 3), random_state_dict(cfg, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["conv_in.weight"], random_state_dict(cfg, 4)["conv_in.weight"])
    net = SeerUNet(sample_size=32, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64)
    assert list(net.state_dict().keys()) == list(a.keys())
    net.load_state_dict(a, strict=True)
    with pytest.raises(RuntimeError):
        bad = dict(a); bad.pop("conv_in.bias")
        net.load_state_dict(bad, strict=True)
    # zero-init proj_out at construction, like the reference (attention.py:126-127)
    fresh = SeerUNet(sample_size=32, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64)
    assert float(fresh.down_blocks[0].attentions[0].proj_out.weight.abs().max()) == 0.0
    assert fresh.config.cross_attention_dim == 64 and fresh.config["layers_per_block"] == 2
    with pytest.raises(RuntimeError, match="CUDA only"):
        fresh(torch.zeros(1, 4, 2, 8, 8), 3, torch.zeros(1, 2, 77, 64))
    # every way of changing the parameters bumps the weights version captured CUDA graphs are keyed on (graph.GraphedUNet.matches)
    v0 = fresh._weights_version
    fresh.load_state_dict(a, strict=True)
    v1 = fresh._weights_version
    fresh.float()
    v2 = fresh._weights_version
    fresh.reset_parameters()
    assert v0 < v1 < v2 < fresh._weights_version and fresh._packed is None


def test_sampler_schedule_matches_reference_golden(golden_dir):
    from seervideoldm_b200.ddim import DDIMSampler
    for S in (30, 10):
        g = torch.load(os.path.join(golden_dir, f"schedule_{S}.pt"), weights_only=False)
        s = DDIMSampler("cpu")
        s.make_schedule(S, verbose=False)
        assert np.array_equal(s.ddim_timesteps, g["timesteps"].numpy())
        assert torch.equal(s.ddim_alphas, g["alphas"])
        assert np.array_equal(np.asarray(s.ddim_alphas_prev), g["alphas_prev"].numpy())
        assert torch.equal(s.ddim_sqrt_one_minus_alphas, g["sqrt_one_minus_alphas"])
        assert torch.equal(s.alphas_cumprod, g["alphas_cumprod_fp32"])
        a, ap = g["alphas"], g["alphas_prev"].float()
        assert torch.equal(s._coef[:, 0], torch.sqrt(1 - a)) and torch.equal(s._coef[:, 1], a.sqrt())
        assert torch.equal(s._coef[:, 2], ap.sqrt()) and torch.equal(s._coef[:, 3], (1 - ap).sqrt())
    assert len(DDIMSampler("cpu").__dict__) > 0


def test_packing_layouts():
    cin = 128
    w = torch.arange(2 * cin * 9, dtype=torch.float32).reshape(2, cin, 3, 3) % 251   # exact in bf16
    p = packing.pack_conv3x3(w).float()
    # K order [Cin/64][ky][kx][64]: one K block per (64-channel slab, tap)
    for co in range(2):
        for ci in range(0, cin, 7):
            for ky in range(3):
                for kx in range(3):
                    assert p[co, (ci // 64) * 9 * 64 + (ky * 3 + kx) * 64 + ci % 64] == w[co, ci, ky, kx]
    with pytest.raises(ValueError):
        packing.pack_conv3x3(torch.zeros(2, 3, 3, 3))
    perm = packing.geglu_permutation(128)
    assert perm[:32].tolist() == list(range(32)) and perm[32:64].tolist() == list(range(128, 160))
    assert perm[64:96].tolist() == list(range(32, 64)) and sorted(perm.tolist()) == list(range(256))
    sc = torch.ones(2, 64, 1, 1)
    assert packing.pack_conv3x3(w, sc).shape == (2, 9 * cin + 64)
    assert packing.pack_conv_out(torch.zeros(4, 8, 3, 3)).shape == (4, 9, 8)


def test_shared_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/seer_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "seer_b200.h")).read()
    declared = set(re.findall(r"\b(seer_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.seer_b200_version().startswith(b"seer_b200")
    assert lib.seer_b200_groupnorm_workspace_floats(2, 12288) == 2 * 384 * 32 * 2
    # argument validation happens before any CUDA call: bad shapes are rejected with -1 even without a device
    assert lib.seer_b200_gemm_bf16(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0, None, 0, None, 0, 0, None) == -1


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under seervideoldm_b200/ may import it, and bench.py only inside the
    cpu baseline / reference-arm function."""
    pkg = os.path.join(ROOT, "seervideoldm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, fn)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), f"{fn} imports oracle"
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef):
            uses = any(isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle") for n in ast.walk(fn))
            assert (not uses) or fn.name == "cpu_eval_fn", fn.name
        else:
            assert not (isinstance(fn, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(fn))


def test_ops_fail_loudly_without_cuda():
    from seervideoldm_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(ValueError, match="no CPU path"):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    with pytest.raises(ValueError, match="no CPU path"):
        ops.layernorm(torch.zeros(4, 320), torch.ones(320), torch.zeros(320))


def test_fstext_module_mirrors_reference_api():
    """FSTextTransformer: constructor / set_numframe / state-dict schema of unet_3d_condition.py:379-398; strict loading;
    CUDA-only forward (no CPU fallback)."""
    import pytest
    import torch
    from seervideoldm_b200 import FSTextTransformer
    from seervideoldm_b200.weights import fstext_schema, random_fstext_state_dict
    m = FSTextTransformer(num_frames=16, num_layers=2)
    sch = fstext_schema(16, 2)
    assert list(m.state_dict().keys()) == list(sch.keys())
    assert all(tuple(v.shape) == sch[k] for k, v in m.state_dict().items())
    assert float(m.learnable_query.abs().sum()) == 0.0 and float(m.pos_embed.abs().sum()) == 0.0    # reference zero-init
    sd = random_fstext_state_dict(16, 2, seed=0)
    m.load_state_dict(sd, strict=True)
    bad = dict(sd)
    bad.pop("norm.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    m.set_numframe(12)
    assert m.num_frames == 12
    assert m.enable_xformers_memory_efficient_attention() is None and m.set_attention_slice("auto") is None
    with pytest.raises(ValueError):
        m.set_precision("fp16")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 77, 768))          # parameters on the CPU: the kernels are CUDA-only, nothing falls back
    with pytest.raises(ValueError):
        m(torch.zeros(1, 77, 512))
    # nearest-neighbour frame resize of the position table (unet_3d_condition.py:477-478)
    q = m._query_tokens(1, 77).view(12, 77, 768)
    idx = (torch.arange(12) * 16 // 12)
    assert torch.equal(q, (sd["learnable_query"] + sd["pos_embed"][:, :, :77])[0, idx])


def test_c_host_compiles_against_the_header_and_links_the_library(tmp_path):
    """The boundary is a C ABI: a plain C99 host includes include/seer_b200.h, links libseer_b200.so and calls the entry
    points that need no GPU (version string, descriptor size, argument validation -> SEER_EINVAL)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib.build(verbose=False)
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "seer_b200.h"
int main(void) {
  SeerGemmDesc d;
  memset(&d, 0, sizeof d);
  if ((int)sizeof d != seer_b200_gemm_desc_size()) { printf("desc size %d vs %d\n", (int)sizeof d, seer_b200_gemm_desc_size()); return 2; }
  int rc = seer_b200_gemm_ex(&d, 0);                 /* all-zero descriptor: rejected before any CUDA call */
  int rc2 = seer_b200_gemm_ex(0, 0);
  printf("%s rc=%d rc2=%d\n", seer_b200_version(), rc, rc2);
  return (rc < 0 && rc2 < 0) ? 0 : 3;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.join(ROOT, "seervideoldm_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lseer_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "rc=-1" in out and "rc2=-1" in out and out.split()[0]


def test_decode_latents_matches_the_reference_expressions():
    """ddim_sampling_utils.py:37-41 restated with einops on a stub VAE: frame-major flattening, 1/0.18215 scale, clamp."""
    from types import SimpleNamespace
    from einops import rearrange
    from seervideoldm_b200.pipeline import ddim_sample, decode_latents

    class StubVAE:                                           # per-image, so frame order mistakes change the result
        def decode(self, z):
            self.seen = z
            img = torch.nn.functional.interpolate(z[:, :3] * 0.3 + z[:, 3:4] * torch.arange(z.shape[0]).reshape(-1, 1, 1, 1) * 0.01,
                                                  scale_factor=8, mode="nearest")
            return SimpleNamespace(sample=img)

    g = torch.Generator().manual_seed(0)
    lat = torch.randn(2, 4, 5, 4, 4, generator=g)
    vae = StubVAE()
    got = decode_latents(vae, lat)
    z = 1 / 0.18215 * rearrange(lat, "n c f h w -> (n f) c h w")
    want = torch.clamp((rearrange(StubVAE().decode(z).sample, "(n f) c h w -> n c f h w", f=5) + 1.0) / 2.0, min=0.0, max=1.0)
    assert torch.equal(vae.seen, z) and got.shape == (2, 3, 5, 32, 32) and torch.equal(got, want)
    assert float(got.min()) >= 0.0 and float(got.max()) <= 1.0 and float((got == 0).float().mean()) > 0     # the clamp is active

    class StubSampler:                                       # ddim_sample = sampler.sample(...) + decode, arguments as the reference passes them
        def sample(self, **kw):
            self.kw = kw
            return lat, {}

    smp = StubSampler()
    out = ddim_sample(smp, "unet", vae, (2, 4, 5, 4, 4), "c", "xT", "x0", ddim_steps=30, scale=1.0, uc="uc")
    assert torch.equal(out, want)
    assert smp.kw["unconditional_conditioning"] is None and smp.kw["S"] == 30 and smp.kw["batch_size"] == 2     # scale 1 drops uc
    assert smp.kw["shape"] == (4, 5, 4, 4) and smp.kw["eta"] == 0.0 and smp.kw["is_3d"] is True and smp.kw["x_T"] == "xT"


def test_upsample_phase_weights_equal_nearest2x_conv():
    """packing.pack_upsample_phases: nearest-2x + conv3x3 (resnet.py:47-61) == four 2x2-tap convs on the low-res image."""
    import torch.nn.functional as F
    from seervideoldm_b200.packing import pack_upsample_phases
    g = torch.Generator().manual_seed(5)
    n, C, Co, H, W = 2, 64, 8, 5, 4
    x = torch.randn(n, C, H, W, generator=g, dtype=torch.double)
    w = torch.randn(Co, C, 3, 3, generator=g)
    w = w.bfloat16().float() * 0.25            # values whose pairwise / 4-way sums stay exactly representable in bf16
    w = (w * 64).round() / 64
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w.double(), padding=1)
    out = torch.empty_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for ph, wp in enumerate(pack_upsample_phases(w)):
        py, px = ph >> 1, ph & 1
        wk = wp.double().reshape(Co, C // 64, 2, 2, 64).permute(0, 1, 4, 2, 3).reshape(Co, C, 2, 2)
        out[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], wk)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)


def test_torch_op_library_registration():
    """the C ABI is registered as a torch custom-op library (`torch.ops.seer_b200.*`): every op has a schema with
    caller-owned mutable outputs, a CUDA kernel only (CPU tensors fail in the dispatcher — no CPU path) and a fake kernel."""
    import torch
    from seervideoldm_b200 import ops, torch_ops  # noqa: F401
    want = {"gemm_ex", "gemm_row_parts", "groupnorm", "groupnorm_from_stats", "layernorm", "attention", "scta_row_index", "rope",
            "timestep_embedding", "small_linear", "conv_in", "conv_out", "upsample2x", "im2col3x3", "cast_bf16", "cfg_ddim_update",
            "split3", "geglu_f32", "rope_table", "rope_inplace", "rope_apply_table", "softmax_rows", "tokens_to_nchw"}
    assert want <= set(torch_ops.OP_NAMES)
    for name in want:
        op = getattr(torch.ops.seer_b200, name).default
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CUDA"), name
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CPU"), name
    schema = str(torch.ops.seer_b200.gemm_ex.default._schema)
    assert "Tensor(a!)? out_f32" in schema and "Tensor(d!)? row_stats_out" in schema and schema.endswith("-> int")
    # a CPU tensor never reaches a kernel: the Python API rejects it, the raw op has no CPU dispatch entry
    with pytest.raises(ValueError):
        ops.cast_bf16(torch.zeros(8))
    with pytest.raises(NotImplementedError):
        torch.ops.seer_b200.cast_bf16(torch.zeros(8), torch.zeros(8, dtype=torch.bfloat16))
    # fake (meta) kernels: the ops trace under FakeTensorMode without launching anything
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        a = torch.empty(256, 64, dtype=torch.bfloat16, device="cuda")
        w = torch.empty(128, 64, dtype=torch.bfloat16, device="cuda")
        out = torch.empty(256, 128, dtype=torch.float32, device="cuda")
        rc = torch.ops.seer_b200.gemm_ex(a, None, None, w, None, 0, None, out, None, False, None, None, None, 0.0, None, 1, [], 0,
                                         None, 0, 0, 0)
        assert rc == 0
        torch.ops.seer_b200.cast_bf16(torch.empty(8, device="cuda"), torch.empty(8, dtype=torch.bfloat16, device="cuda"))


def test_vae_schema_and_oracle_cpu():
    """The AutoencoderKL state-dict schema is the published SD-1.5 VAE's (83 653 863 parameters, diffusers key names) and the CPU
    oracle runs both directions on it; known-answer: encode/decode shapes and the x8 geometry."""
    import math
    from oracle import vae_oracle as vo
    from seervideoldm_b200.vae import AutoencoderKL, random_vae_state_dict, vae_schema
    schema = vae_schema()
    assert sum(math.prod(v) for v in schema.values()) == 83_653_863 and len(schema) == 248
    for k in ("encoder.down_blocks.1.resnets.0.conv_shortcut.weight", "decoder.up_blocks.2.resnets.0.conv_shortcut.weight",
              "decoder.mid_block.attentions.0.proj_attn.bias", "encoder.down_blocks.2.downsamplers.0.conv.weight",
              "decoder.up_blocks.0.upsamplers.0.conv.bias", "quant_conv.weight", "post_quant_conv.bias"):
        assert k in schema, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in schema and "decoder.up_blocks.3.upsamplers.0.conv.weight" not in schema
    sd = random_vae_state_dict(seed=0)
    vae = AutoencoderKL()
    vae.load_state_dict(sd, strict=True)                      # same keys / shapes as the module tree
    assert list(vae.state_dict().keys()) == list(schema.keys())
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    img = vo.decode(sd, z)
    assert img.shape == (1, 3, 64, 64)
    m = vo.encode_moments(sd, img)
    assert m.shape == (1, 8, 8, 8)
    with pytest.raises(RuntimeError):
        vae.decode(z)                                         # CPU module: no CPU path
